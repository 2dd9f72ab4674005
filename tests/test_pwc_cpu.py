"""CPU: eavsr_b200.pwc.PWCNET reproduces the reference PWCNET's parameter names and shapes (so a
`pwc-default` state dict loads strictly) -- checked against the key list recorded from the reference itself
in tests/golden/pwc_net.npz (make_golden.py gen_pwc)."""
import os

import numpy as np
import pytest
import torch

from eavsr_b200.pwc import PWCNET, estimate

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_state_dict_keys_and_shapes_equal_the_reference():
    g = np.load(os.path.join(GOLD, "pwc_net.npz"))
    sd = PWCNET().state_dict()
    keys = sorted(sd)
    assert keys == list(g["keys"])
    assert [str(tuple(sd[k].shape)) for k in keys] == list(g["key_shapes"])
    # the reference renames 'module*' -> 'net*' when loading sniklaus' blob (models/pwc_net.py:249-251)
    renamed = {k.replace("net", "module", 1): v for k, v in sd.items()}
    PWCNET().load_state_dict({k.replace("module", "net"): v for k, v in renamed.items()}, strict=True)


def test_cpu_tensors_raise():
    net = PWCNET().eval()
    with pytest.raises(NotImplementedError):        # the cost volume has no CPU path (correlation.py:324-325)
        estimate(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64), net)
