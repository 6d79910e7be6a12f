"""The packed time-major schedule of the recurrent path (host logic, runs on CPU)."""
import numpy as np
import torch

from nnacousticmodeling_b200.recurrent_engine import Schedule


def test_schedule_covers_every_utterance_step_once():
    rng = np.random.default_rng(0)
    steps = rng.integers(1, 40, size=77)
    s = Schedule(steps, 32, 2, 4, torch.device("cpu"))
    assert s.n_rows == steps.sum() and s.n_batches == 3
    utt = s.order[s.row_sorted_utt]
    seen = set(zip(utt.tolist(), s.row_step.tolist()))
    assert len(seen) == s.n_rows
    assert seen == {(u, t) for u in range(77) for t in range(steps[u])}
    # device tables reproduce row(t, u) = row0 + base[t] + u
    row0, boff = s.d_row0.numpy(), s.d_boff.numpy()
    base, ulen = s.d_base.numpy(), s.d_utt_len.numpy()
    for r in rng.integers(0, s.n_rows, size=200):
        su, t = int(s.row_sorted_utt[r]), int(s.row_step[r])
        b, u = divmod(su, 32)
        assert row0[b] + base[boff[b] + t] + u == r
        assert ulen[su] == steps[s.order[su]] > t
    # every (batch, direction) item is scheduled exactly once, groups are balanced by LPT
    items = sorted(zip(s.d_item_batch.tolist(), s.d_item_dir.tolist()))
    assert items == [(b, d) for b in range(3) for d in range(2)]
    gs = s.d_group_start.tolist()
    assert gs[0] == 0 and gs[-1] == 6 and len(gs) == s.n_groups + 1


def test_schedule_lanes_with_streams():
    """Several streams per CTA group: lanes = groups x streams, every lane's items are direction-sorted (a group
    works on one direction at a time), and the critical path accounts for the per-group direction switch."""
    rng = np.random.default_rng(1)
    steps = rng.integers(5, 60, size=300)
    s = Schedule(steps, 16, 2, 3, torch.device("cpu"), streams=4)
    assert s.n_batches == 19 and s.n_groups == 3 and s.n_lanes == 12
    gs = s.d_group_start.tolist()
    assert len(gs) == s.n_lanes + 1 and gs[-1] == s.n_items == 38
    dirs, batches = s.d_item_dir.tolist(), s.d_item_batch.tolist()
    for ln in range(s.n_lanes):
        d = dirs[gs[ln]:gs[ln + 1]]
        assert d == sorted(d)
    assert sorted(zip(batches, dirs)) == [(b, d) for b in range(19) for d in range(2)]
    bsteps = s.d_steps.numpy()
    load = np.zeros((12, 2), np.int64)
    for ln in range(12):
        for i in range(gs[ln], gs[ln + 1]):
            load[ln, dirs[i]] += bsteps[batches[i]]
    assert s.max_group_steps == load.reshape(3, 4, 2).max(axis=1).sum(axis=1).max()
    assert s.d_counters.numel() == 12
    # fewer batches than lanes: no empty groups are launched
    s = Schedule([9] * 20, 16, 1, 9, torch.device("cpu"), streams=4)
    assert s.n_batches == 2 and s.n_groups == 1 and s.n_lanes == 4


def test_schedule_single_and_equal_lengths():
    s = Schedule([5], 32, 1, 9, torch.device("cpu"))
    assert s.n_rows == 5 and s.n_groups == 1 and s.d_base.tolist() == [0, 1, 2, 3, 4, 5]
    s = Schedule([1] * 40, 32, 1, 9, torch.device("cpu"))
    assert s.n_rows == 40 and np.array_equal(s.order, np.arange(40))  # stable sort keeps input order
    assert s.d_row0.tolist() == [0, 32]


def test_pair_assignment_covers_every_item_once_and_costs_solo_tail():
    """Two-stream groups (the 128-slot kernels): every (batch, direction) item lands on exactly one lane, lanes are
    direction-sorted, and the reported critical path follows  solo * longer lane + (1 - solo) * shorter lane."""
    from nnacousticmodeling_b200.recurrent_engine import assign_lanes
    rng = np.random.default_rng(3)
    for n_dirs in (1, 2):
        for n_batches in (1, 2, 7, 29, 60):
            bsteps = np.sort(rng.integers(5, 800, n_batches))[::-1]
            for ratio in (1.0, 0.8):
                per_lane, n_groups, crit = assign_lanes(bsteps, n_dirs, 9, 2, ratio)
                assert len(per_lane) == 2 * n_groups and 1 <= n_groups <= 9
                flat = [it for ln in per_lane for it in ln]
                assert sorted(flat) == sorted((b, d) for b in range(n_batches) for d in range(n_dirs))
                worst = 0.0
                for g in range(n_groups):
                    cost = 0.0
                    for d in range(n_dirs):
                        a, b = (sum(int(bsteps[i]) for i, dd in per_lane[2 * g + s] if dd == d) for s in range(2))
                        cost += ratio * max(a, b) + (1 - ratio) * min(a, b)
                    worst = max(worst, cost)
                    for s in range(2):
                        dirs = [dd for _, dd in per_lane[2 * g + s]]
                        assert dirs == sorted(dirs)
                assert abs(crit - np.ceil(worst)) <= 1
                # never worse than the trivial bounds
                assert crit >= ratio * int(bsteps[0]) - 1
    # a lone long batch is cheaper next to an idle sibling than the all-busy figure
    _, _, busy = assign_lanes([785, 100, 90], 1, 9, 2, 1.0)
    _, _, solo = assign_lanes([785, 100, 90], 1, 9, 2, 0.8)
    assert solo < busy


def test_mixed_schedule_packs_every_frame_once():
    """MixedSchedule (long utterances in the 32-slot kernel, the rest in the 128-slot kernel): one packed row space,
    every (utterance, step) exactly once, part B's rows start where part A's end."""
    from nnacousticmodeling_b200.recurrent_engine import MixedSchedule
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 300, 500)
    ms = MixedSchedule(lens, 64, 2, 32, 1, 128, 7, 2, 0.8, 2, torch.device("cpu"))
    (sa, nb_a), (sb, nb_b) = ms.parts
    assert (nb_a, nb_b) == (32, 128) and ms.n_rows == int(lens.sum()) == sa.n_rows + sb.n_rows
    assert sorted(ms.order.tolist()) == list(range(500))
    assert int(sa.d_row0[0]) == 0 and int(sb.d_row0[0]) == sa.n_rows
    utt = ms.order[ms.row_sorted_utt]
    pairs = set(zip(utt.tolist(), ms.row_step.tolist()))
    assert len(pairs) == ms.n_rows
    assert np.all(ms.row_step < lens[utt])
    # the long part holds exactly the 64 longest utterances
    assert sorted(lens[ms.order[:64]].tolist(), reverse=True) == sorted(lens.tolist(), reverse=True)[:64]
    assert sa.n_groups <= 2 and sb.n_groups <= 7


def test_phase_split_covers_every_utterance_once(monkeypatch):
    """Two-subset forward: every utterance in exactly one subset, the first holds the shortest ones (about 30 % of the
    frames), small shards stay whole, NNAM_RNN_PHASES=1 switches it off."""
    from nnacousticmodeling_b200.recurrent_engine import _phase_split
    rng = np.random.default_rng(5)
    lens = rng.integers(100, 800, 1344)
    parts = _phase_split(lens)
    assert len(parts) == 2
    both = np.concatenate(parts)
    assert sorted(both.tolist()) == list(range(1344))
    assert all(np.all(np.diff(p) > 0) for p in parts)
    assert lens[parts[0]].max() <= lens[parts[1]].min()
    frac = lens[parts[0]].sum() / lens.sum()
    assert 0.2 < frac < 0.4
    assert len(_phase_split(lens[:200])) == 1          # few utterances
    assert len(_phase_split(np.full(400, 100))) == 1   # few frames
    monkeypatch.setenv("NNAM_RNN_PHASES", "1")
    assert len(_phase_split(lens)) == 1


def test_cost_model_gives_long_groups_several_batches(monkeypatch):
    """pick_schedule on a test-shaped set with the kernels' measured step costs (faked here: the C ABI's nnam_rnn_plan
    needs a device): the mixed schedule wins, its long part holds MORE batches than groups (a group works through its
    batches one after the other while the group with the longest batch walks that one), and the packed row space
    still covers every (utterance, step) exactly once."""
    import types
    from nnacousticmodeling_b200 import recurrent_engine as R
    from nnacousticmodeling_b200 import synth
    plans = {16: (0, 9, 9000, 4), 32: (0, 7, 4300, 1), 64: (0, 9, 7900, 1), 128: (0, 9, 12500, 2)}
    monkeypatch.setattr(R.ops, "rnn_plan", lambda cell, hidden, nb, nsplit: plans[nb])
    monkeypatch.setattr(R.ops, "rnn_solo_step_cycles", lambda cell, hidden, nb, nsplit: 10100)
    monkeypatch.delenv("NNAM_RNN_MIXED", raising=False)
    monkeypatch.delenv("NNAM_RNN_MIXED_GROUPS", raising=False)
    monkeypatch.delenv("NNAM_RNN_MIXED_BATCHES", raising=False)
    plan = types.SimpleNamespace(cell=R.CELL_LSTM, hidden=512, n_dirs=1, split=False)
    steps = synth.synth_lengths(np.random.default_rng(1234), 1344) + 5
    sched, tag = R.pick_schedule(plan, steps, torch.device("cpu"))
    assert isinstance(sched, R.MixedSchedule) and tag[0] == "mixed"
    _, k_long, g_long, nb_long = tag
    assert nb_long == 32 and k_long % 32 == 0 and k_long // 32 > g_long >= 2   # several batches per long group
    (sa, _), (sb, nb_b) = sched.parts
    assert nb_b == 128 and sa.n_groups == g_long and sa.n_groups + sb.n_groups <= 9
    assert sched.n_rows == int(steps.sum())
    utt = sched.order[sched.row_sorted_utt]
    assert len(set(zip(utt.tolist(), sched.row_step.tolist()))) == sched.n_rows
    # the long part's critical lane is not longer than the bulk part's modelled time (that is what the search balances)
    assert sa.max_group_steps * 4300 <= 1.15 * sb.max_group_steps * 12500
    # forcing one batch per group reproduces the previous behaviour
    monkeypatch.setenv("NNAM_RNN_MIXED_GROUPS", "3")
    monkeypatch.setenv("NNAM_RNN_MIXED_BATCHES", "3")
    plan2 = types.SimpleNamespace(cell=R.CELL_LSTM, hidden=512, n_dirs=1, split=False)
    _, tag2 = R.pick_schedule(plan2, steps, torch.device("cpu"))
    assert tag2 == ("mixed", 96, 3, 32)


def test_fused_output_layer_gate(monkeypatch):
    """engine.fused_head_ok: one net, plain head, 512..2048 classes, fan-in <= 1024; environment overrides."""
    import types
    from nnacousticmodeling_b200 import engine
    monkeypatch.delenv("NNAM_FUSED_HEAD", raising=False)
    net = types.SimpleNamespace(n_out=1909, network="ff")
    plain = engine.HeadSpec()
    assert engine.fused_head_ok([net], plain, 512) and engine.fused_head_ok([net], plain, 1024)
    assert not engine.fused_head_ok([net], plain, 2048)
    assert not engine.fused_head_ok([net, net], plain, 512)
    assert not engine.fused_head_ok([net], engine.HeadSpec(rpl={"W": 1}), 512)
    assert not engine.fused_head_ok([net], engine.HeadSpec(final_normalize=False), 512)
    assert not engine.fused_head_ok([net], engine.HeadSpec(weights=[0.5]), 512)
    assert engine.fused_head_ok([net], engine.HeadSpec(prior=np.zeros(1909, np.float32), prior_scale=0.7), 512)
    assert not engine.fused_head_ok([types.SimpleNamespace(n_out=39, network="ff")], plain, 128)
    assert not engine.fused_head_ok([types.SimpleNamespace(n_out=1909, network="tdnn")], plain, 512)
    assert not engine.fused_head_ok([types.SimpleNamespace(n_out=4000, network="ff")], plain, 512)
    monkeypatch.setenv("NNAM_FUSED_HEAD", "0")
    assert not engine.fused_head_ok([net], plain, 512)
    monkeypatch.setenv("NNAM_FUSED_HEAD", "force")
    assert engine.fused_head_ok([net], plain, 2048)
    assert engine.fused_head_ok([types.SimpleNamespace(n_out=39, network="ff")], plain, 128)
