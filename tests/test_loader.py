"""CPU tests of the loader / trainer host logic (SURVEY 8b loader boundary, 8a a11): id arithmetic, sampling order,
priority table, chunking of the training loop.  No device work: the scene store only queues host arrays."""
import numpy as np
import pytest

from dynamics_b200 import t7
from dynamics_b200.loader import DataSampler, GeneralDataSampler, PrioritySampler
from dynamics_b200.train import Logger, chunk_len, parse_args, parse_lua_list
from oracle import batcher, scenes


class _NoDevice:
    """Stands in for NPEModel: the samplers only touch it when a batch is selected."""


def _write_dataset(root, name, n_balls, n_scenes, seed, sets=("trainset", "valset", "testset")):
    st = scenes.raw_to_state(scenes.gen_balls(n_scenes, n_balls, 60, seed=seed))
    foc, ctx, _, _ = batcher.expand_for_each_object(st)
    cuts = batcher.cut_batches(foc, ctx)
    per = max(1, len(cuts) // len(sets))
    files = {}
    for k, s in enumerate(sets):
        d = root / name / "batches" / s
        d.mkdir(parents=True)
        files[s] = cuts[k * per:(k + 1) * per]
        for i, (f, c) in enumerate(files[s], 1):
            t7.save_batch_file(str(d / f"batch{i}"), f, c)
    return files


def _args(root, folders, **kw):
    a = {"data_root": str(root), "dataset_folders": folders, "maxwinsize": 60, "winsize": 3, "num_past": 2,
         "num_future": 1, "relative": True, "sim": False, "subdivide": True, "shuffle": True, "model": _NoDevice(),
         "batch_size": 50, "rng": np.random.default_rng(7)}
    a.update(kw)
    return a


def test_subbatch_ids_follow_reference(tmp_path):
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    a = _args(tmp_path, ["balls_n3_t60_ex100_m"]); a["dataset_folder"] = str(tmp_path / "balls_n3_t60_ex100_m")
    d = DataSampler.create("trainset", a)
    assert d.num_batches == 2 and d.num_subbatches_per_batch == 20 and d.total_batches == 40
    for sid in range(1, 41):
        b, o = d.subbatch_location(sid)
        ob, oo = batcher.subbatch_id_to_file_offset(sid)
        assert (b - 1, o - 1) == (ob, oo)
    assert sorted(d.batch_idxs.tolist()) == list(range(1, 41))           # torch.randperm(total_batches)
    seen = []
    for _ in range(40):                                                  # one epoch, without replacement
        batch = d.sample_sequential_batch()
        seen.append(d.current_sampled_id)
        b, o = d.subbatch_location(d.current_sampled_id)
        assert np.array_equal(batch.scene, np.arange(50) + d.file_base[b - 1]) and np.all(batch.t0 == o - 1)
        assert np.all(batch.t0 + 3 <= 22)                                # only frames 1..22 are ever trained on
    assert sorted(seen) == list(range(1, 41))
    d.sample_sequential_batch()
    assert d.current_batch == 1                                          # wraps (data_sampler.lua:134)


def test_priority_sampler_protocol(tmp_path):
    ps = PrioritySampler(4, np.random.default_rng(0))
    for i, w in enumerate([0.5, 0.1, 2.0], 1):
        ps.update_batch_weight(i, w)
    assert not ps.table_is_full and ps.get_hardest_batch() == [2.0, 3]
    with pytest.raises(AssertionError):
        ps.set_alpha(0.0); ps.sample(1)                                  # normalize() asserts min > 0
    ps.update_batch_weight(4, 1e-3)
    assert ps.table_is_full
    ps.set_alpha(0.0)
    draws = np.bincount([ps.sample(2) for _ in range(4000)], minlength=5)[1:]
    p = np.array([0.5, 0.1, 2.0, 1e-3]) ** 2
    assert np.allclose(draws / 4000, p / p.sum(), atol=0.03)


def test_general_sampler_cycles_and_weights(tmp_path):
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    _write_dataset(tmp_path, "balls_n4_t60_ex100_m", 4, 100, seed=2)
    g = GeneralDataSampler.create("trainset", _args(tmp_path, ["balls_n3_t60_ex100_m", "balls_n4_t60_ex100_m"]))
    assert g.num_batches == 40 + 40 and len(g.datasamplers) == 2         # 2 files x 20 offsets per folder
    # without modulo: the first dataset is exhausted before the second starts (general_data_sampler.lua:127-131)
    order = [g.sample_sequential_batch(False)[1] for _ in range(80)]
    assert order == [1] * 40 + [2] * 40
    assert [g.sample_sequential_batch(True)[1] for _ in range(4)] == [1, 2, 1, 2]
    # priority protocol: sequential until every batch of the drawn dataset has a weight, then multinomial
    for _ in range(400):
        batch, ds = g.sample_priority_batch(1)
        g.update_batch_weight(0.25 + 0.001 * g.current_sampled_id)
    assert all(d.priority_sampler.table_is_full for d in g.datasamplers) and g.has_seen_all_batches
    loss, sid, ds = g.get_hardest_batch()
    assert sid == 40 and abs(loss - 0.29) < 1e-12 and ds == g.current_dataset
    batch, ds = g.sample_random_batch()
    assert len(batch) == 50 and 1 <= ds <= 2
    # example scenes of different object counts share one store
    assert g.datasamplers[0].store is g.datasamplers[1].store and g.datasamplers[1].file_base[0] == 100


def test_chunks_end_where_the_reference_acts():
    periods = [7, 20, 5]
    done, fired = 0, []
    while done < 60:
        n = chunk_len(done, periods, 60 - done)
        done += n
        fired.append(done)
    expect = sorted({k for k in range(1, 61) if any(k % p == 0 for p in periods)})
    assert fired == expect


def test_cli_mirrors_main_lua(tmp_path):
    mp = parse_args(["-model", "npe", "-nbrhd", "-rs", "-opt", "adam", "-name", "{'balls'}", "-dataset_folders",
                     "{'balls_n3_t60_ex50000_m','balls_n4_t60_ex50000_m'}", "-test_dataset_folders", "{'balls_n6_t60_ex50000_m'}"])
    assert mp.dataset_folders == ["balls_n3_t60_ex50000_m", "balls_n4_t60_ex50000_m"] and mp.name == "balls"
    assert (mp.lr, mp.lrdecay, mp.lrdecay_every, mp.lrdecayafter, mp.batch_size, mp.max_iter) == (3e-4, 0.99, 2500, 50000, 50, 1200000)
    assert mp.winsize == 3 and mp.savedir == "logs/balls" and mp.rs and mp.nbrhd
    assert parse_lua_list('{"a", "b"}') == ["a", "b"]
    dbg = parse_args(["-debug"])
    assert (dbg.batch_size, dbg.max_iter, dbg.print_every, dbg.val_every) == (5, 60, 1, 20)
    lg = Logger(str(tmp_path / "x" / "train.log"))
    lg.add({"a": 1.5, "b": 2}); lg.add({"a": 0.25, "b": 3})
    assert open(tmp_path / "x" / "train.log").read().splitlines() == ["a\tb\t", "1.5\t2\t", "0.25\t3\t"]
