"""GPU: the tcgen05 implicit-GEMM conv against the CUDA-core direct conv on every layer shape class
of car.onnx / armor.onnx (k, stride, Cin, Cout, map size, batch), through the C ABI."""
import pytest

import rm_radar_b200 as rr

pytestmark = pytest.mark.gpu

# (n, h, w, cin, cout, k, stride, act, residual, out_f32)
SHAPES = [
    (1, 160, 160, 64, 64, 3, 1, 1, 0, 0),     # car head 160^2 3x3
    (1, 160, 160, 32, 32, 3, 1, 1, 1, 0),     # C2f(64) bottleneck, BK=32 / SWIZZLE_64B, shortcut
    (1, 320, 320, 32, 64, 3, 2, 1, 0, 0),     # backbone layer 1, stride 2
    (1, 160, 160, 64, 128, 3, 2, 1, 0, 0),    # layer 3 stride 2
    (3, 160, 160, 64, 128, 3, 2, 1, 0, 0),    # same, three images: the four parity-class patches must not bleed across images
    (2, 80, 80, 128, 256, 3, 2, 1, 0, 0),     # stride 2 with two channel chunks per patch
    (1, 80, 80, 128, 128, 1, 1, 1, 0, 0),     # 1x1
    (1, 40, 40, 768, 256, 1, 1, 1, 0, 0),     # neck C2f cv1, K=768
    (1, 20, 20, 256, 256, 3, 1, 1, 1, 0),     # 20^2: ragged tiles (TW=4..)
    (1, 20, 20, 1024, 512, 1, 1, 1, 0, 0),    # SPPF cv2
    (1, 40, 40, 256, 256, 3, 2, 1, 0, 0),     # stride 2 small map
    (1, 160, 160, 64, 64, 1, 1, 0, 0, 1),     # head final box conv, no act, fp32 out
    (1, 160, 160, 64, 1, 1, 1, 0, 0, 1),      # head final cls conv (Cout=1 -> padded 16), fp32 out
    (3, 80, 80, 128, 12, 1, 1, 0, 0, 1),      # armor cls final, batch 3
    (8, 10, 10, 512, 512, 3, 1, 1, 1, 0),     # armor 10^2 maps: tiles span images (TN>1)
    (5, 20, 20, 384, 384, 3, 2, 1, 0, 0),     # armor stride-2 with batch (parity view must not bleed)
    (2, 40, 40, 96, 192, 1, 1, 1, 0, 0),      # Cin=96 (BK=32 x3), Cout=192 (2 x 96)
    (2, 20, 20, 576, 384, 1, 1, 1, 0, 0),     # Cin=576
    (20, 80, 80, 128, 128, 3, 1, 1, 0, 0),    # max batch
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "n{}_{}x{}_c{}-{}_k{}s{}_a{}r{}f{}".format(*s))
def test_conv_umma_matches_direct(shape):
    n, h, w, cin, cout, k, s, act, res, f32 = shape
    diff, ref, _ = rr.conv_selftest(n, h, w, cin, cout, k, s, act, res, f32, seed=hash(shape) & 0xffff)
    # both paths: fp16 operands, fp32 accumulate; they differ in summation order and (fp16 out) one rounding
    tol = (2e-3 if f32 else 6e-3) * max(1.0, ref)
    assert diff == diff and diff <= tol, f"max|diff| {diff} > {tol} (max ref {ref})"
