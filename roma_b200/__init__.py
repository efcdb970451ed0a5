"""roma_b200 — B200-native implementation of RoMa's dense `match()` / `sample()` path.

Drop-in for `romatch`'s public surface on that path (`romatch/__init__.py:2`):

    from roma_b200 import roma_outdoor
    model = roma_outdoor(device="cuda", weights=..., dinov2_weights=...)
    warp, certainty = model.match(im_A, im_B)
    matches, conf = model.sample(warp[0], certainty[0])
"""
DEBUG_MODE = False
RANK = 0
GLOBAL_STEP = 0
STEP_SIZE = 1
LOCAL_RANK = -1


def __getattr__(name):
    # factories import torch + the CUDA library lazily so that `import roma_b200.arch` stays light
    if name in ("roma_outdoor", "roma_indoor", "tiny_roma_v1_outdoor", "roma_model"):
        from . import model_zoo
        return getattr(model_zoo, name)
    raise AttributeError(name)
