from .binding import LtransLib, Params, LTGPU_F32, LTGPU_F64, LtransError  # noqa: F401
