"""CPU: oracle/buffer.py against the committed golden fixture (written by the unmodified reference
buffer, oracle/gen_golden_buffer.py): ring writes, Gumbel-top-k index sets, adjust with kills."""
import torch

from golden_util import load_fixture
from oracle.buffer import OracleBuffer, topk_set


def replay(make_buffer, add, sample, adjust, state_of):
    fx = load_fixture("prioritised_buffer")
    cfg = fx["config"]
    buf = make_buffer(cfg)
    it = iter(fx["batches"])
    for st in fx["steps"]:
        if st["op"] == "init":
            for _ in range(st["n_batches"]):
                add(buf, *next(it))
        elif st["op"] == "add":
            add(buf, *next(it))
        elif st["op"] == "sample":
            idx = sample(buf, st["k"], st["z"])
            assert torch.equal(torch.sort(idx.cpu()).values, st["indices_sorted"]), "sampled index set"
            continue
        elif st["op"] == "adjust":
            adjust(buf, st["adj"], st["log_q"], st["indices"])
        got, want = state_of(buf), st["state"]
        for k in ("x", "log_w", "log_q_old"):
            assert torch.equal(got[k].cpu(), want[k]), f"{st['op']}: {k}"      # bit-exact (copies / one add)
        for k in ("current_index", "is_full", "can_sample"):
            assert got[k] == want[k], f"{st['op']}: {k}"


def test_oracle_buffer_matches_reference_fixture():
    replay(lambda c: OracleBuffer(c["dim"], c["max_length"], c["min_sample_length"]),
           lambda b, x, lw, lq: b.add(x, lw, lq),
           lambda b, k, z: b.sample(k, z=z)[3],
           lambda b, adj, lq, idx: b.adjust(adj, lq, idx),
           lambda b: dict(x=b.x, log_w=b.log_w, log_q_old=b.log_q_old, current_index=b.current_index,
                          is_full=b.is_full, can_sample=b.can_sample))


def test_topk_set_properties():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1000, generator=g) * 4
    logits[5] = -float("inf")
    z = torch.zeros(1000)
    idx = topk_set(logits, z, 10)
    assert torch.equal(idx, torch.sort(torch.argsort(logits, descending=True)[:10]).values)
    assert 5 not in idx.tolist()
