"""GPU test (-m gpu): the paired / shared fp64 divisions of the device code (ibis_b200/csrc/scalar.h: div_pair,
div_shared, div_z -- the compiler's own fast-path sequence run for two quotients at once) give the quotients of the
plain IEEE division, bit for bit, on 10^8 operand pairs spanning the whole exponent range (zeros, denormals,
infinities, NaNs, equal operands included).  The reference divides plainly (x86-64 divsd): this is what keeps the
flux calculators and the limiter bit-identical to it."""
import pytest

import cases
from ibis_b200.lib import Context

pytestmark = pytest.mark.gpu


def test_paired_divisions_equal_the_ieee_division():
    import torch
    assert torch.cuda.is_available()
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    ctx = Context(g, config)
    for seed in (1, 2026):
        assert ctx.debug_division_check(50_000_000, seed) == 0
    ctx.close()
