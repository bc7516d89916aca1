"""GPU: detection heads.  pv_roi_align_fwd vs torchvision.ops.roi_align goldens, and the
reference's slow_r50_detection / slowfast_r50_detection (trunk + ResNetRoIHead, models/head.py:394-482,
net.py:47-74) vs goldens produced by the real reference (tests/golden/detection.pt, oracle/gen_golden.py).

Tolerances: fp32 storage - rtol 1e-3 / atol 1e-4 * scale (north star); f16 tensor-core path - measured bounds on
f16-grid weights / clips like tests/test_gpu_models.py (box coordinates stay fp32 on both paths)."""
import os

import pytest
import torch

from pytorchvideo_b200 import testing as TS
import pytorchvideo_b200.models.hub as PH

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "detection.pt"), weights_only=False)


@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_roi_align_matches_torchvision_golden(dtype):
    from pytorchvideo_b200 import ops
    x, boxes, settings = TS.roi_align_case()
    for (osz, scale, sr), ref in zip(settings, GOLD["roi_align"]["outputs"]):
        got = ops.roi_align(x.cuda(), boxes.cuda(), osz, scale, sr, dtype).cpu()
        assert got.shape == ref.shape
        err = (got - ref).abs()
        if dtype == "f32":
            assert float(err.max()) <= 2e-6 * max(1.0, float(ref.abs().max())), (osz, sr, float(err.max()))
        else:   # the feature map is exactly representable in f16; one rounding of the stored result
            assert bool((err <= 1e-3 * ref.abs() + 1e-6).all()), (osz, sr, float(err.max()))


# case: (min in-band fraction, max |d|/max|ref|) on the f16 path - measured on B200 (profiles/r02_parity.md)
F16_BOUNDS = {
    "slow_r50_detection": (0.86, 8e-4),                  # 0.900 / 5.4e-4 (logits)
    "slowfast_r50_detection": (0.95, 6e-4),              # 0.983 / 3.2e-4 (logits)
    "slow_r50_detection_sigmoid": (0.88, 5e-3),          # 0.925 / 3.2e-3 (probabilities: |d p| <= |d logit| / 4)
}


@pytest.mark.parametrize("precision", ["f32", "f16"])
@pytest.mark.parametrize("case", sorted(TS.DETECTION_CASES))
def test_detection_model_matches_reference_golden(case, precision):
    from pytorchvideo_b200 import config
    g = GOLD[case]
    model, inp, boxes, is_sf = TS.build_detection_case(case, PH)
    assert abs(TS.state_checksum(model) - g["state_checksum"]) <= 1e-6 * abs(g["state_checksum"])
    ref = g["output"]
    config.set_precision(precision)
    try:
        model.cuda()
        x = [t.cuda() for t in inp] if is_sf else inp.cuda()
        out = model(x, boxes.cuda()).float().cpu()
        out2 = model(x, boxes.cuda()).float().cpu()                 # cached plan + graph replay
        # a different number of boxes compiles its own plan; the shared boxes give the same rows
        out_k = model(x, boxes[:3].cuda()).float().cpu()
    finally:
        config.set_precision("f16")
        model.cpu()
    assert out.shape == ref.shape and torch.equal(out, out2)
    scale = max(1.0, float(ref.abs().max()))
    err = (out - ref).abs()
    inside = float((err <= 1e-3 * ref.abs() + 1e-4 * scale).float().mean())
    rel = float(err.max()) / scale
    print("PARITY %s %s: max|d|/max|ref| = %.3e, fraction within rtol1e-3/atol1e-4 = %.3f" % (case, precision, rel, inside))
    if precision == "f32":
        assert bool((err <= 1e-3 * ref.abs() + 1e-4 * scale).all()), "max err %.3e (scale %.3g)" % (float(err.max()), scale)
    else:
        lo, hi = F16_BOUNDS[case]
        assert rel <= hi and inside >= lo, (rel, inside)
    assert out_k.shape == (3, ref.shape[1])
    assert torch.allclose(out_k, out[:3], rtol=1e-3, atol=1e-3 * scale)


def test_detection_rejects_bad_boxes():
    model, inp, boxes, _ = TS.build_detection_case("slow_r50_detection_sigmoid", PH)
    model.cuda()
    with pytest.raises(RuntimeError):
        model(inp.cuda(), torch.zeros(2, 6, device="cuda"))          # rotated-box format
    with pytest.raises(RuntimeError):
        model(inp.cuda(), boxes)                                     # CPU boxes: no silent host path
    model.cpu()
