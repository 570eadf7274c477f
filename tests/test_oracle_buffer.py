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


# ---- fab/utils/replay_buffer.py (un-prioritised buffer) ------------------------------------------
def replay_uniform(make_buffer, add, sample, state_of):
    """Walk tests/golden/replay_buffer.pt (written by the unmodified reference ReplayBuffer,
    oracle/gen_golden_buffer.py: main_uniform): ring writes with add counters, rank-weighted multinomial
    without replacement (= exponential race on the recorded variates) at temperature 1 and 0."""
    fx = load_fixture("replay_buffer")
    cfg = fx["config"]
    for run in fx["runs"]:
        buf = make_buffer(cfg, run["temperature"])
        it = iter(run["batches"])
        for st in run["steps"]:
            if st["op"] == "init":
                for _ in range(st["n_batches"]):
                    add(buf, *next(it), init=True)
            elif st["op"] == "add":
                add(buf, *next(it), init=False)
            elif st["op"] == "sample":
                x, log_w, idx = sample(buf, st["k"], st["q"])
                assert torch.equal(idx.cpu(), st["indices"]), "sampled indices (reference order)"
                assert torch.equal(x.cpu(), st["x"]) and torch.equal(log_w.cpu(), st["log_w"])
                continue
            got, want = state_of(buf), st["state"]
            for k in ("x", "log_w", "add_count"):
                assert torch.equal(got[k].cpu(), want[k]), f"{st['op']}: {k}"
            for k in ("current_index", "current_add_count", "is_full", "can_sample"):
                assert got[k] == want[k], f"{st['op']}: {k}"


def _init_add(b, x, lw, init):
    b.add(x, lw)
    if init:                                  # replay_buffer.py:52-57: the fill loop resets the add counter
        b.current_add_count = 0 if not b.can_sample else 1


def test_oracle_replay_buffer_matches_reference_fixture():
    from oracle.buffer import OracleReplayBuffer

    def sample(b, k, q):
        x, lw, idx = b.sample(k, q=q)
        return x, lw, idx
    replay_uniform(lambda c, T: OracleReplayBuffer(c["dim"], c["max_length"], c["min_sample_length"], temperature=T),
                   _init_add, sample,
                   lambda b: dict(x=b.x, log_w=b.log_w, add_count=b.add_count, current_index=b.current_index,
                                  current_add_count=b.current_add_count, is_full=b.is_full, can_sample=b.can_sample))


def test_multinomial_without_replacement_is_the_exponential_race():
    """What the restatement (and the device buffer) rely on: torch.multinomial(replacement=False) on the CPU
    draws one Exponential(1) variate per entry and returns topk(probs / q) in descending order."""
    from oracle.buffer import exponential_like, race_topk
    for n, k, T in [(300, 64, 1.0), (500, 192, 0.5), (200, 50, 0.0), (1000, 999, 2.0)]:
        rank = torch.randint(1, 9, (n,), generator=torch.Generator().manual_seed(n)).float()
        probs = torch.pow(1 / rank, T)
        torch.manual_seed(7)
        idx = torch.multinomial(probs, num_samples=k, replacement=False)
        torch.manual_seed(7)
        assert torch.equal(idx, race_topk(probs, exponential_like(probs), k))
