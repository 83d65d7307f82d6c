"""Transport.training_losses mirror (forward-only values) against fixtures the imported reference
produced (oracle/gen_golden_train_data.py: transport/transport.py:112-215, path.py:114-136) and, in the build container,
against the live reference with the CPU oracle as the model."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
import va_vae_b200.transport as T


def stand_in_model(x, t, y=None):
    return torch.tanh(x * (0.5 + t.view(-1, 1, 1, 1))) + 0.01 * y.view(-1, 1, 1, 1).float()


CASES = torch.load(os.path.join(GOLDEN_DIR, "training_losses.pt"), weights_only=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_training_losses_match_reference_fixture(name):
    case = CASES[name]
    transport = T.create_transport("Linear", "velocity", None, None, None, **case["kw"])
    torch.manual_seed(11)
    np.random.seed(12)
    terms = transport.training_losses(stand_in_model, case["x1"], dict(y=case["y"]),
                                      shifted_mu=0.3 if case["kw"].get("shift_lg") else 0)
    assert set(terms) == set(case["terms"])
    for k, want in case["terms"].items():
        assert terms[k].shape == want.shape
        assert torch.allclose(terms[k], want, rtol=1e-6, atol=1e-7), k


def test_noise_and_score_parameterisations_are_refused():
    tr = T.create_transport("Linear", "noise", None, None, None)
    with pytest.raises(NotImplementedError):
        tr.training_losses(stand_in_model, torch.zeros(2, 4, 8, 8), dict(y=torch.zeros(2, dtype=torch.long)))


def test_against_live_reference_with_the_oracle_model():
    from oracle import ref_shim
    from oracle import lightningdit_oracle as O
    if not ref_shim.reference_available():
        pytest.skip("/root/reference only exists in the build container")
    _, tr = ref_shim.import_reference()
    cfg = O.DiTConfig(input_size=8, patch_size=1, in_channels=4, hidden_size=64, depth=2, num_heads=2, num_classes=10)
    sd = O.init_state_dict(cfg, seed=1)
    model = lambda x, t, y=None: O.forward(cfg, sd, x, t, y)
    x1 = torch.randn(3, 4, 8, 8, generator=torch.Generator().manual_seed(2))
    y = torch.tensor([1, 5, 9])
    out = []
    for mod in (tr, T):
        transport = mod.create_transport("Linear", "velocity", None, None, None, use_cosine_loss=True, use_lognorm=True)
        torch.manual_seed(3)
        np.random.seed(4)
        out.append(transport.training_losses(model, x1, dict(y=y)))
    for k in out[0]:
        assert torch.equal(out[0][k], out[1][k]), k
