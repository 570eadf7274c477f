import os

import torch

from oracle.realnvp import OracleRealNVP, randomize_last_layers
from oracle.targets import OracleGMM, OracleManyWell

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rebuild_flow(fx) -> OracleRealNVP:
    """Flow weights are regenerated from the seed (oracle/gen_golden.py:build_flow) and verified
    against the checksum stored with the fixture."""
    import numpy as np
    cfg = fx["config"]
    torch.manual_seed(cfg["flow_seed"])
    f = OracleRealNVP(cfg["dim"], cfg["K"], cfg["npd"])
    if cfg["K"]:
        randomize_last_layers(f, 0.05, seed=cfg["flow_seed"] + 1)
    if cfg["base_scale"] is not None:
        with torch.no_grad():
            f._nf_model.q0.log_scale.fill_(float(np.log(cfg["base_scale"])))
    chk = float(sum(p.detach().double().abs().sum() for p in f.state_dict().values()))
    assert abs(chk - fx["flow_checksum"]) <= 1e-9 * abs(chk), "flow weights differ from fixture"
    return f


def rebuild_target(fx):
    cfg = fx["config"]
    t = cfg["target"]
    if t[0] == "mw":
        return OracleManyWell(cfg["dim"])
    torch.manual_seed(0)
    return OracleGMM(cfg["dim"], t[1], t[2], t[3])
