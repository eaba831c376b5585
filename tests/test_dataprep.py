"""CPU tests of the data-format rows: json <-> states <-> batch files, against the oracle codec."""
import json

import numpy as np

from dynamics_b200 import dataprep, t7
from oracle import batcher, scenes


def _json_from_raw(raw):
    return {"trajectories": dataprep.data2table(raw), "config": {}}


def test_json_to_states_matches_oracle_codec(tmp_path):
    raw = scenes.gen_walls(3, 6, seed=4, variant="I")
    p = tmp_path / "walls-n2-t6-ex3-wI-chksize3-0.json"
    json.dump(_json_from_raw(raw), open(p, "w"))
    got = dataprep.raw_to_states(dataprep.load_data_json(json.load(open(p))))
    assert np.array_equal(got, scenes.raw_to_state(raw))
    assert dataprep.dataset_flags("data/walls_n2_t60_ex50000_wU/jsons")["num_obj"] == 33
    assert dataprep.dataset_flags("balls_n4_t60_ex200_m") == {"env": "balls", "n": 4, "t": 60, "ex": 200, "num_obj": 4}
    assert dataprep.global_params_from_name("x/tower_n5_t120_ex100/jsons/a.json")[0] == 1
    assert dataprep.global_params_from_name("balls_n3_t60_ex10_m/jsons/a.json") == (0, 0, 0)


def test_record_trajectories_roundtrip(tmp_path):
    raw = scenes.gen_balls(4, 3, 8, seed=2)
    st = scenes.raw_to_state(raw)
    out = dataprep.record_trajectories(st, str(tmp_path / "pred.json"), config={"env": "balls"})
    back = dataprep.load_data_json(json.load(open(tmp_path / "pred.json")))
    assert np.allclose(back[..., :6], raw[..., :6], atol=2e-3) and np.array_equal(back[..., 6:9], raw[..., 6:9])
    assert out["trajectories"][0][0][0]["objtype"] == "ball"


def test_batch_files_follow_reference_layout(tmp_path):
    raws = [scenes.gen_balls(100, 3, 60, seed=s) for s in (1, 2)]
    files = []
    for k, r in enumerate(raws):
        p = tmp_path / f"chunk{k}.json"
        json.dump(_json_from_raw(r), open(p, "w")); files.append(str(p))
    counts = dataprep.json_to_batch_files(files, str(tmp_path / "out"))
    assert sum(counts.values()) == 12 and counts == {"trainset": 10, "valset": 1, "testset": 1}   # 600 examples / 50
    f, c = t7.load_batch_file(str(tmp_path / "out" / "batches" / "trainset" / "batch1"))
    assert f.shape == (50, 60, 22) and c.shape == (50, 2, 60, 22)
    st = scenes.raw_to_state(raws[0])
    of, oc, _, _ = batcher.expand_for_each_object(st)
    df, dc = dataprep.expand_for_each_object(st)
    assert np.array_equal(of, df) and np.array_equal(oc, dc)      # focus-index-major expansion == the oracle's
    sc, n_obj = dataprep.scenes_from_json(files)
    assert sc.shape == (200, 3, 60, 22) and np.all(n_obj == 3)
