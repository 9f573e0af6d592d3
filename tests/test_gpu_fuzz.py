"""A short run of the randomised differential test (tools/fuzz_parity.py): random small bilayers,
every kernel family against the oracle port."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_bilayers_match_the_port(seed):
    import fuzz_parity
    rng = np.random.default_rng(seed)
    for k in range(12):
        desc, bad = fuzz_parity.one_case(rng, k)
        assert not bad, (desc, bad)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_random_windows_and_batches_give_the_resident_outputs(seed):
    """Random systems, every native analysis, random window sizes / host->device batch sizes:
    streamed and batch-wise runs must reproduce the single-batch resident run exactly."""
    from pybilt_b200 import BilayerAnalyzer, synthetic as syn
    from tests import cases
    rng = np.random.default_rng(seed)
    n = int(rng.choice([24, 64, 150]))
    spec = syn.BilayerSpec(n, cases.CG3, seed=int(rng.integers(1, 1 << 30)), npt=float(rng.choice([0.0, 2e-3])),
                           leaflet_z=float(rng.choice([1.5, 20.0])), z_step=float(rng.choice([0.15, 0.7])),
                           step=float(rng.choice([0.5, 6.0])))
    F = int(rng.integers(9, 30))
    top, traj = syn.generate(spec, F)
    analyses = ["msd_multi mm n_tau 4 n_sigma 3", "com_lateral_rdf r1 resname_1 POPC resname_2 DOPE",
                "com_lateral_rdf r2 resname_1 all resname_2 all", "nnf n1 resname_1 POPC resname_2 DOPE n_neighbors 3",
                "nnf n2 resname_1 DOPE resname_2 DOPE n_neighbors 3", "apl_grid apl", "thickness_grid bt",
                "thickness_leaflet_distance tld", "halperin_nelson hn", "ald ald_1", "flip_flop ff", "apl_box ab",
                "acm acm_1", "ac ac_1", "vcm v1", "area_fluctuation af"]

    def run(window, batch_frames, incremental):
        ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
        for a in analyses:
            ba.add_analysis(a)
        ba.rep_settings["lipid_grid"]["n_xbins"] = ba.rep_settings["lipid_grid"]["n_ybins"] = 6
        ba.window_frames, ba.batch_frames, ba._incremental = window, batch_frames, incremental
        ba.run_analysis()
        return cases.flatten_outputs({k: ba.get_analysis_data(k) for k in ba.get_analysis_ids()})

    want = run(None, F, False)
    variants = [(None, int(rng.integers(2, F)), True), (int(rng.integers(4, F)), None, True),
                (int(rng.integers(4, F)), int(rng.integers(2, 6)), True)]
    for window, bf, inc in variants:
        got = run(window, bf, inc)
        assert sorted(got) == sorted(want)
        for k in want:
            assert np.array_equal(np.asarray(got[k]), np.asarray(want[k]), equal_nan=True), (seed, window, bf, k)
