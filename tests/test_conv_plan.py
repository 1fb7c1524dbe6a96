"""Launch plans of the tcgen05 conv (`rmr_conv_plan`, host only): every conv shape of the two networks at the batch
sizes the cascade uses must get a plan that fits the hardware — 148 SMs, 224 KB of shared memory per CTA, 512 TMEM
columns — and covers every output tile exactly once."""
import ctypes as C

import pytest

from rm_radar_b200 import _lib, engine
from tests import fixtures as fx

SM = 148
SMEM_MAX = 224 * 1024

# (h, w, cin, cout, k, stride) of a few layers of both graphs, used when the ONNX files are not in the snapshot
FALLBACK = [(320, 320, 32, 64, 3, 2), (160, 160, 64, 64, 1, 1), (160, 160, 32, 32, 3, 1), (80, 80, 128, 256, 3, 2),
            (40, 40, 512, 256, 1, 1), (20, 20, 256, 256, 3, 1), (20, 20, 1024, 512, 1, 1), (160, 160, 64, 1, 1, 1),
            (80, 80, 128, 12, 1, 1), (10, 10, 512, 128, 3, 1), (20, 20, 384, 384, 3, 2)]


def conv_shapes():
    if not fx.have_onnx():
        return FALLBACK
    shapes = set()
    for name in ("car", "armor"):
        plan = engine.compile_onnx(fx.onnx(name))
        for op in plan.ops:
            if op.type == engine.OP_CONV and op.cin_pad != 4:      # the stem has its own kernel
                shapes.add((op.src.H, op.src.W, op.src.C, op.dst.C, op.k, op.stride))
    return sorted(shapes)


@pytest.mark.parametrize("batch", [1, 2, 7, 16, 20])
def test_every_layer_gets_a_plan_that_fits(batch):
    lib = _lib.load()
    shapes = conv_shapes()
    assert len(shapes) >= len(FALLBACK)
    for (h, w, cin, cout, k, s) in shapes:
        out = (C.c_int * 16)()
        _lib.check(lib.rmr_conv_plan(batch, h, w, cin, cout, k, s, out))
        (ver, n, splits, halo, mtiles, ctas, tiles_per_cta, kb_per_tile, sa, sb, resident, smem, tw, th, tn, bk) = list(out)
        where = f"batch {batch} {h}x{w} c{cin}->{cout} k{k}s{s}: {list(out)}"
        assert ver == 2, where                                    # the round-2 kernel takes every tensor-core layer
        assert n % 16 == 0 and 16 <= n <= 256, where              # UMMA N, one accumulator <= 256 TMEM columns (two buffers)
        assert splits == 1, where                                 # split-K is gone
        assert tw * th * tn == 128, where                         # UMMA M
        assert bk in (16, 32, 64) and (cin % bk == 0 or bk == 16), where
        assert 0 < smem <= SMEM_MAX, where
        assert 1 <= ctas <= SM, where
        assert sa >= 1 and sb >= 0, where
        if halo:
            assert k == 3 and s == halo, where                    # 1: one 18 x pw patch (stride 1); 2: four parity-class patches (stride 2)
        ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
        want_mtiles = -(-wo // tw) * -(-ho // th) * -(-batch // tn)
        assert mtiles == want_mtiles, where
        n_tiles = -(-max(cout, 16) // n)
        # every (pixel tile, channel tile) pair has a CTA slot
        assert ctas * tiles_per_cta >= mtiles * n_tiles or ctas * tiles_per_cta >= mtiles, where
        assert kb_per_tile >= 1, where


def test_plan_rejects_nonsense():
    lib = _lib.load()
    out = (C.c_int * 16)()
    assert lib.rmr_conv_plan(0, 20, 20, 64, 64, 3, 1, out) != 0
    assert lib.rmr_conv_plan(1, 20, 20, 64, 64, 5, 1, out) != 0
