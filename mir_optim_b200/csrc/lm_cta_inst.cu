// lm_cta_inst.cu -- model dispatch of the general batched LM kernel (lm_cta.cuh), both precisions.
#include "lm_cta.cuh"
#include "runtime.cuh"

namespace mirb200 {

template <class T>
int launch_cta_model(const mir_model_desc& model, size_t n, const typename Num<T>::Settings& st, const SmallBatchArgs& args, cudaStream_t stream)
{
    if (model.model >= (uint32_t)MIR_MODEL_USER_BASE) return launch_user_model<T>(model, n, st, args, stream);
    switch (model.model) {
    case MIR_MODEL_EXPDECAY2: return launch_cta<CtaFromLarge<LModelExpDecay2<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_EXPTAU3:   return launch_cta<CtaFromLarge<LModelExpTau3<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_EXPDECAY3: return launch_cta<CtaFromLarge<LModelExpDecay3<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_GAUSS4:    return launch_cta<CtaFromLarge<LModelGauss4<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_SUMEXP:    return launch_cta<CtaFromLarge<LModelSumExp<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_GAUSSMIX:  return launch_cta<CtaFromLarge<LModelGaussMix<T>, T>, T>(st, args, n, model, stream);
    case MIR_MODEL_SPLINE:
        if (!model.aux) { set_error("mir_optim_b200: MIR_MODEL_SPLINE needs the knots in model->aux"); return MIR_B200_EINVAL; }
        return launch_cta<CtaSpline<T>, T>(st, args, n, model, stream);
    default:
        set_error("mir_optim_b200: model id not available in the batched path");
        return MIR_B200_EUNSUPPORTED;
    }
}
template int launch_cta_model<double>(const mir_model_desc&, size_t, const Num<double>::Settings&, const SmallBatchArgs&, cudaStream_t);
template int launch_cta_model<float>(const mir_model_desc&, size_t, const Num<float>::Settings&, const SmallBatchArgs&, cudaStream_t);

}  // namespace mirb200
