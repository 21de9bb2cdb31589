"""Process-wide settings of the sm_100a path."""
import os

# default: the fp32-parity tensor-core mode (validated against the oracle at <= 1e-4)
_precision = os.environ.get("SINNERF_B200_PRECISION", "f16x3")


def set_precision(name: str) -> None:
    """Arithmetic of the field MLP: 'fp32' (FFMA, exact), 'f16x3' / 'bf16x3' (tcgen05, split operands,
    fp32-parity), 'bf16' (tcgen05 single pass).  Everything outside the MLP is always fp32."""
    from . import _lib
    _lib.precision_id(name)
    global _precision
    _precision = name


def get_precision() -> str:
    return _precision
