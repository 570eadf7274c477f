"""Pin oracle/buffer.py against the UNMODIFIED reference buffer and write
tests/golden/prioritised_buffer.pt.  TEST INFRASTRUCTURE (build container only):

    python -m oracle.gen_golden_buffer

Scenario (fab/utils/prioritised_replay_buffer.py): fill by `initial_sampler` up to the minimum
length, three more `add`s that wrap the ring, `sample_n_batches` (Gumbel-top-k without replacement;
the Gumbel draw is recorded by re-seeding), `adjust` with finite and non-finite entries, another
sample.  The fixture holds every input, the reference's sampled index sets and its buffer state
after every step.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.buffer import (OracleBuffer, gumbel_like, topk_set, OracleReplayBuffer,   # noqa: E402
                           exponential_like, race_topk)
from oracle.ref_loader import load_reference                       # noqa: E402


def main():
    load_reference()
    from fab.utils.prioritised_replay_buffer import PrioritisedReplayBuffer as RefBuffer
    dim, max_length, min_len, B = 6, 500, 128, 96
    g = torch.Generator().manual_seed(42)
    batches = [(torch.randn(B, dim, generator=g), torch.randn(B, generator=g) * 3, torch.randn(B, generator=g))
               for _ in range(7)]
    it = iter(batches)
    ref = RefBuffer(dim, max_length, min_len, lambda: next(it), device="cpu")
    orc = OracleBuffer(dim, max_length, min_len)
    n_init = 0
    for b in batches:
        if orc.can_sample:
            break
        orc.add(*b)
        n_init += 1
    steps = []

    def same_state():
        return (torch.equal(ref.buffer.x, orc.x) and torch.equal(ref.buffer.log_w, orc.log_w)
                and torch.equal(ref.buffer.log_q_old, orc.log_q_old)
                and ref.current_index == orc.current_index and ref.is_full == orc.is_full
                and ref.can_sample == orc.can_sample)

    def snap():
        return dict(x=orc.x.clone(), log_w=orc.log_w.clone(), log_q_old=orc.log_q_old.clone(),
                    current_index=orc.current_index, is_full=orc.is_full, can_sample=orc.can_sample)

    assert same_state(), "init fill differs"
    steps.append(dict(op="init", n_batches=n_init, state=snap()))
    for b in batches[n_init:n_init + 4]:                      # wraps the ring (2*96 + 4*96 > 500)
        ref.add(*b)
        orc.add(*b)
        assert same_state(), "add differs"
        steps.append(dict(op="add", state=snap()))
    for rnd in range(2):
        k = 64 * 3
        torch.manual_seed(100 + rnd)
        max_index = ref.max_length if ref.is_full else ref.current_index
        z = gumbel_like(ref.buffer.log_w[:max_index])          # what the reference will draw
        torch.manual_seed(100 + rnd)
        data = ref.sample_n_batches(64, 3)
        ref_idx = torch.cat([d[3] for d in data])
        x_o, lw_o, lq_o, idx_o = orc.sample(k, z=z)
        assert torch.equal(torch.sort(ref_idx).values, idx_o), "sampled index set differs"
        assert torch.equal(torch.cat([d[0] for d in data]), ref.buffer.x[ref_idx])
        steps.append(dict(op="sample", k=k, z=z, max_index=max_index, indices_sorted=idx_o.clone(),
                          ref_indices=ref_idx.clone()))
        gg = torch.Generator().manual_seed(7 + rnd)
        adj = torch.randn(k, generator=gg)
        lq = torch.randn(k, generator=gg)
        adj[3] = float("nan"); adj[10] = float("inf"); lq[20] = float("-inf")
        ref.adjust(adj, lq, ref_idx)
        orc.adjust(adj, lq, ref_idx)
        assert same_state(), "adjust differs"
        steps.append(dict(op="adjust", adj=adj, log_q=lq, indices=ref_idx.clone(), state=snap()))
    out = dict(config=dict(dim=dim, max_length=max_length, min_sample_length=min_len, batch=B),
               batches=batches, steps=steps)
    path = os.path.join(ROOT, "tests", "golden", "prioritised_buffer.pt")
    torch.save(out, path)
    print(f"oracle buffer == reference buffer on every step; wrote {path} "
          f"({os.path.getsize(path) / 1e3:.0f} kB)")


def main_uniform():
    """fab/utils/replay_buffer.py (the buffer of fab/train_with_buffer.py): fill by `initial_sampler`, adds that
    wrap the ring, `sample_n_batches` at temperature 1 (the default), 0 (uniform) -- the exponential variates of
    torch.multinomial are recorded by re-seeding -- more adds, another sample.  Writes tests/golden/replay_buffer.pt."""
    load_reference()
    from fab.utils.replay_buffer import ReplayBuffer as RefBuffer
    dim, max_length, min_len, B = 5, 400, 150, 64
    out = dict(config=dict(dim=dim, max_length=max_length, min_sample_length=min_len, batch=B), runs=[])
    for temperature in (1.0, 0.0):
        g = torch.Generator().manual_seed(11)
        batches = [(torch.randn(B, dim, generator=g), torch.randn(B, generator=g) * 2) for _ in range(14)]
        it = iter(batches)
        ref = RefBuffer(dim, max_length, min_len, lambda: next(it), device="cpu", temperature=temperature)
        orc = OracleReplayBuffer(dim, max_length, min_len, temperature=temperature)
        it2 = iter(batches)
        orc.fill(lambda: next(it2))
        n_init = sum(1 for _ in range(len(batches))) - sum(1 for _ in it2)

        def same_state():
            return (torch.equal(ref.buffer.x, orc.x) and torch.equal(ref.buffer.log_w, orc.log_w)
                    and torch.equal(ref.buffer.add_count, orc.add_count) and ref.current_index == orc.current_index
                    and ref.current_add_count == orc.current_add_count and ref.is_full == orc.is_full
                    and ref.can_sample == orc.can_sample)

        def snap():
            return dict(x=orc.x.clone(), log_w=orc.log_w.clone(), add_count=orc.add_count.clone(),
                        current_index=orc.current_index, current_add_count=orc.current_add_count,
                        is_full=orc.is_full, can_sample=orc.can_sample)

        assert same_state(), "init fill differs"
        steps = [dict(op="init", n_batches=n_init, state=snap())]
        k_next = n_init
        for rnd in range(3):
            for _ in range(3):                                   # 3 x 64 rows per round: wraps in round 2
                b = batches[k_next]; k_next += 1
                ref.add(*b); orc.add(*b)
                assert same_state(), "add differs"
                steps.append(dict(op="add", state=snap()))
            k = 32 * 3
            torch.manual_seed(200 + rnd)
            q = exponential_like(orc.probs())                    # what torch.multinomial will draw
            torch.manual_seed(200 + rnd)
            data = ref.sample_n_batches(32, 3)
            x_r, lw_r = torch.cat([d[0] for d in data]), torch.cat([d[1] for d in data])
            x_o, lw_o, idx_o = orc.sample(k, q=q)
            assert torch.equal(x_r, x_o) and torch.equal(lw_r, lw_o), "sampled rows differ"
            steps.append(dict(op="sample", k=k, q=q, indices=idx_o.clone(), x=x_r.clone(), log_w=lw_r.clone()))
        out["runs"].append(dict(temperature=temperature, batches=batches, steps=steps))
    path = os.path.join(ROOT, "tests", "golden", "replay_buffer.pt")
    torch.save(out, path)
    print(f"oracle replay buffer == reference ReplayBuffer on every step (temperature 1 and 0); wrote {path} "
          f"({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
    main_uniform()
