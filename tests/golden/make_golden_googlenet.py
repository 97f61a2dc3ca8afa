"""Golden features of the frame-feature extractor, from the REAL torchvision GoogLeNet wrapped the way the reference
wraps it (src/helpers/video_helper.py:36-40,61-73): run here (torchvision 0.26 is in this image; there is no network,
so the weights are the seeded synthetic ones of oracle.googlenet_oracle under torchvision's own state-dict names).

    python tests/golden/make_golden_googlenet.py        ->  tests/golden/googlenet.npz
"""
import hashlib
import os
import sys

import numpy as np
import torch
from numpy import linalg
from torch import nn
from torchvision import models

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import googlenet_oracle as gno  # noqa: E402

W_SEED, X_SEED, N_FRAMES = 4242, 777, 3


def main():
    torch.set_num_threads(4)
    net = models.googlenet(weights=None, aux_logits=False, init_weights=False)
    params = gno.synth_googlenet_params(W_SEED)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("fc.") or k.endswith("num_batches_tracked") for k in missing), missing
    # video_helper.py:37-40 (CPU instead of .cuda())
    model = nn.Sequential(*list(net.children())[:-2]).eval()
    x = gno.synth_frames(N_FRAMES, X_SEED)
    feats = []
    with torch.no_grad():
        for i in range(N_FRAMES):
            # video_helper.py:61-73: one frame per call, flatten, divide by the norm
            feat = model(x[i:i + 1]).view(-1).cpu().numpy()
            assert feat.shape == (1024,)
            feat /= linalg.norm(feat) + 1e-10
            feats.append(feat)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "googlenet.npz")
    np.savez_compressed(out, feats=np.stack(feats), w_seed=W_SEED, x_seed=X_SEED,
                        x_sha=hashlib.sha256(x.numpy().tobytes()).hexdigest(),
                        torchvision=str(__import__("torchvision").__version__))
    print(out, np.stack(feats).shape)


if __name__ == "__main__":
    main()
