// todo_stubs.cu -- entry points declared in the header whose kernels are not written yet.
// They fail loudly (never compute on the CPU).  Each stub disappears as its kernel lands.
#include "runtime.cuh"
using namespace mirb200;
static int nyi(const char* what) { set_error(std::string("mir_optim_b200: ") + what + " is not implemented yet"); return MIR_B200_EUNSUPPORTED; }
extern "C" {
mir_least_squares_result_d mir_optimize_least_squares_d(const mir_least_squares_settings_d*, size_t, size_t, double*, const double*, const double*,
    mir_slice_d, mir_slice_i, void*, mir_ls_function_d, void*, mir_ls_jacobian_d, void*, mir_ls_thread_manager)
{ nyi("mir_optimize_least_squares_d"); mir_least_squares_result_d r{mir_ls_numericError, 0, 0, 0, __builtin_huge_val(), 0}; return r; }
mir_least_squares_result_s mir_optimize_least_squares_s(const mir_least_squares_settings_s*, size_t, size_t, float*, const float*, const float*,
    mir_slice_s, mir_slice_i, void*, mir_ls_function_s, void*, mir_ls_jacobian_s, void*, mir_ls_thread_manager)
{ nyi("mir_optimize_least_squares_s"); mir_least_squares_result_s r{mir_ls_numericError, 0, 0, 0, __builtin_huge_valf(), 0}; return r; }
void mir_b200_device_model_d(void*, size_t, size_t, const double*, double*) {}
void mir_b200_device_model_jac_d(void*, size_t, size_t, const double*, double*) {}
void mir_b200_device_model_s(void*, size_t, size_t, const float*, float*) {}
void mir_b200_device_model_jac_s(void*, size_t, size_t, const float*, float*) {}
int mir_optimize_least_squares_sharded_d(const mir_least_squares_settings_d*, const mir_model_desc*, size_t, size_t, double*, const double*, const double*, void*, void*, mir_least_squares_result_d*, mir_batch_stats*) { return nyi("mir_optimize_least_squares_sharded_d"); }
int mir_b200_nccl_unique_id(void*) { return nyi("nccl"); }
int mir_b200_nccl_comm_init(void**, int, const void*, int) { return nyi("nccl"); }
int mir_b200_nccl_comm_destroy(void*) { return nyi("nccl"); }
}
