"""CPU: autograd through the oracle restatement against gradients produced by the reference's own modules
(oracle/make_golden_train.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2


@pytest.mark.parametrize("name", ["train_c64_l4", "train_c256_l3"])
def test_oracle_gradients_match_reference(name):
    from oracle import edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    g = load_golden(name)
    C, layers, cycle, B, L, seed, stride = (int(v) for v in g["cfg"])
    sd = {k: v.clone().requires_grad_(True) for k, v in make_wavenet_state_dict(C, layers, seed).items()}
    loss = edm.dsm_loss(torch.from_numpy(g["x"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["sigmas"]),
                        owav.make_net_fn(sd, cycle), 0.2)
    assert torch.allclose(loss.detach(), torch.from_numpy(g["loss"]), rtol=1e-5)
    loss.mean().backward()
    grads = [v.grad.reshape(-1) for v in sd.values()]
    flat = torch.cat(grads)
    assert rel_l2(flat[::stride], g["grad_sub"]) < 1e-5
    norms = np.array([float(x.double().norm()) for x in grads])
    assert np.allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)
