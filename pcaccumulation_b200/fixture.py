"""Deterministic, non-degenerate weights for parity tests and benchmarks.

No released checkpoint is available offline, and the reference's default initialisation is
degenerate for parity work (``ResnetBlockFC.fc_1`` is zero-initialised,
``models/pillar_encoder.py:44``, and every pillar is classified foreground so the ego-motion head has
no background pillars to register).  ``fixture_state_dict`` therefore derives every tensor from
``(seed, key name, shape)`` alone, so the reference model (in the build container) and this package
(on the GPU box) can be loaded with bit-identical weights without shipping a 44 MB checkpoint.
The two segmentation-head output biases are shifted by fixed calibrated constants so that roughly
20 % of the pillars are foreground and roughly 10 % of the foreground points are dynamic on the
synthetic scenes of ``synth.py``.
"""
import zlib

import numpy as np
import torch

# calibrated once against the reference forward on synthetic C1 scenes (oracle/make_golden.py --calibrate)
FB_BIAS_SHIFT = 10.0
MOS_BIAS_SHIFT = 5.0


def _rng(seed, key):
    return np.random.default_rng((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF)


def fixture_state_dict(template, seed=42, fb_shift=None, mos_shift=None):
    """``template``: a state_dict (names -> tensors) giving names, shapes and dtypes."""
    fb_shift = FB_BIAS_SHIFT if fb_shift is None else fb_shift
    mos_shift = MOS_BIAS_SHIFT if mos_shift is None else mos_shift
    out = {}
    for key in sorted(template.keys()):
        ref = template[key]
        shape = tuple(ref.shape)
        g = _rng(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            val = np.array(1, dtype=np.int64)
        elif leaf == "running_mean":
            val = g.normal(0.0, 0.1, shape)
        elif leaf == "running_var":
            val = g.uniform(0.5, 1.5, shape)
        elif key.endswith("ego_motion_head.alpha") or key.endswith("ego_motion_head.beta"):
            val = np.array(-5.0)
        elif leaf == "weight" and len(shape) == 1:  # BatchNorm scale
            val = g.uniform(0.8, 1.2, shape)
        elif leaf == "weight":
            if "upconv" in key:  # ConvTranspose2d [Cin, Cout, 2, 2]: each output sees Cin inputs
                fan_in = shape[0]
            else:
                fan_in = int(np.prod(shape[1:]))
            val = g.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        elif leaf == "bias":
            val = g.normal(0.0, 0.05, shape)
        else:
            raise KeyError(f"fixture: unexpected tensor {key}")
        out[key] = torch.tensor(np.asarray(val), dtype=ref.dtype).reshape(shape)
    k_fb = "semseg_head.seg_head.3.bias"
    k_mos = "motionhead.mos_seg.seg_head.3.bias"
    out[k_fb][0] += fb_shift
    out[k_mos][0] += mos_shift
    return out
