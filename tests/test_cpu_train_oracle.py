"""Training-pass oracle pinned against the reference's own `DefaultSegmentorV2.forward` in train mode followed by
`loss.backward()` (tests/golden/train.npz, made by tests/golden/make_golden_train.py): the loss and the gradient of EVERY backbone
parameter.  This is the target the (not yet built) backward kernels will be tested against."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN
from helpers import replay, t
from oracle import ptv3_oracle as O
from oracle import wrapper_oracle as W
from oracle.weights import synth_state_dict


def test_training_pass_loss_and_gradients_vs_reference():
    Z = np.load(os.path.join(GOLDEN, "train.npz"))
    J = json.load(open(os.path.join(GOLDEN, "train.json")))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in synth_state_dict(J["shapes"]).items()}
    inp = dict(coord=t(Z["coord"]), grid_coord=t(Z["grid_coord"]).long(), offset=t(Z["offset"]), feat=t(Z["feat"]), segment=t(Z["segment"]))
    abar = W.diffusion_hyperparams("cosine", 0, 1000, 1000)[2].float()
    O.BN_TRAIN = True
    try:
        loss, parts = W.training_loss(sd, J["cfg"], inp, t(Z["ts"]), t(Z["noise"]), abar, "GLS", 2, -1, "noise", perm_fn=replay(Z["perms"]))
        loss.backward()
    finally:
        O.BN_TRAIN = False
    assert abs(float(loss.detach()) - float(Z["loss"])) < 2e-5 * float(Z["loss"])
    norms = dict(zip(J["grad_names"], Z["grad_norms"]))
    for name, ref in norms.items():                               # every parameter's gradient norm
        g = sd[name].grad
        assert g is not None, name
        # (biases that feed a train-mode BatchNorm have a mathematically zero gradient: 1e-8 of rounding noise on both sides)
        assert abs(float(g.norm()) - ref) < 2e-3 * ref + 1e-6, (name, float(g.norm()), ref)
    for key in Z.files:
        if key.startswith("grad__"):                              # and a sample of full gradient tensors, element-wise
            g, ref = sd[key[6:]].grad.numpy(), Z[key]
            assert np.abs(g - ref).max() < 1e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, key
    # running statistics take no gradient; batch statistics were used: the eval-mode loss differs
    loss_eval, _ = W.training_loss({k: v.detach() for k, v in sd.items()}, J["cfg"], inp, t(Z["ts"]), t(Z["noise"]), abar, "GLS", 2, -1,
                                   "noise", perm_fn=replay(Z["perms"]))
    assert abs(float(loss_eval) - float(Z["loss"])) > 1e-3
