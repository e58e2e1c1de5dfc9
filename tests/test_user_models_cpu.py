"""Run-time compiled residual models (mir_b200_model_compile): what can be checked without a GPU -- NVRTC runs on the host,
so a model is validated at registration and a broken one is refused with the compiler's log."""
import pytest

import user_models


@pytest.fixture(scope="module")
def eng():
    import mir_optim_b200
    return mir_optim_b200.engine


def test_model_compiles_at_registration_and_ids_are_distinct(eng):
    a = eng.compile_model(user_models.LOGISTIC)
    assert a >= 0x1000
    eng.release_model(a)
    with pytest.raises(Exception):
        eng.release_model(a)                      # already released


def test_compile_error_is_reported_with_the_users_line(eng):
    from mir_optim_b200.engine import B200Error
    bad = user_models.LOGISTIC.replace("exp(-p[1]", "expp(-q[1]")
    with pytest.raises(B200Error) as ei:
        eng.compile_model(bad)
    msg = str(ei.value)
    assert "user_model.cu" in msg and ("expp" in msg or "q" in msg), msg
