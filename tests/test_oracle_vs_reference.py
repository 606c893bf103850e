"""The committed fixtures and the oracle against the LIVE, unmodified reference (build container only: skipped wherever
/root/reference is absent, e.g. on the GPU box).  tests/test_oracle_golden.py pins the oracle to tests/golden/; this file
re-derives a sample of those fixtures from the reference's own modules, so a stale or hand-edited fixture cannot go unnoticed."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mg():
    refimport.activate()
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import make_golden
    return make_golden


@pytest.mark.parametrize("name", ["qm9", "community_small", "zinc250k_v2"])
def test_known_answer_fixtures_come_from_the_reference_modules(mg, name):
    """models/ScoreNetwork_X.py / ScoreNetwork_A.py forward on the SURVEY 8c recipe: the committed kat_<name>.npz is what the
    reference's own nn.Modules return today (bit for bit), and the oracle restatement is within 1e-5 of it."""
    from oracle import gdss_oracle as O
    mx, ma, px, pa = mg.ref_models(name)
    z = np.load(os.path.join(GOLD, f"kat_{name}.npz"))
    x, adj, flags = O.kat_inputs(int(z["N"]), int(z["F"]), int(z["B"]))
    with torch.no_grad():
        sx, sa = mx(x, adj, flags), ma(x, adj, flags)
        ox = O.run_model({k: v for k, v in mx.state_dict().items()}, px, x, adj, flags)
        oa = O.run_model({k: v for k, v in ma.state_dict().items()}, pa, x, adj, flags)
    assert np.array_equal(sx.numpy(), z["score_x"]) and np.array_equal(sa.numpy(), z["score_adj"])
    for ours, ref in ((ox, sx), (oa, sa)):
        assert float((ours - ref).norm() / ref.norm()) < 1e-5


@pytest.mark.parametrize("case_name", ["qm9_s4_n30", "community_subvp_rev_lang_n10", "qm9_rev_lang_ns2_n10"])
def test_trajectory_fixtures_come_from_the_reference_samplers(mg, case_name):
    """solver.py get_pc_sampler / S4_solver run live with the recorded draw stream: the committed traj_<case>.npz final state
    is reproduced (same torch build: bit for bit; otherwise to 1e-5)."""
    from oracle.cases import SHORT_CASES
    case = next(c for c in SHORT_CASES if c["name"] == case_name)
    z = np.load(os.path.join(GOLD, f"traj_{case_name}.npz"))
    assert json.loads(str(z["case"])) == json.loads(json.dumps(case))
    flags, x, adj, n_evals = mg.run_reference_case(case)
    assert int(n_evals) == int(z["n_evals"]) and np.array_equal(flags.numpy(), z["flags"])
    for live, fix in ((x.numpy(), z["x"]), (adj.numpy(), z["adj"])):
        assert np.allclose(live, fix, rtol=1e-5, atol=1e-6)


def test_nspdk_vectoriser_matches_the_reference_live(mg):
    """evaluation/eden.py `vectorize(complexity=4, discrete=True)` on integer-labelled molecule graphs of the C3 reference run
    against oracle/stats.py `nspdk_matrix`: the same sparse feature matrix."""
    from evaluation.eden import vectorize
    from oracle import stats as S
    z = np.load(os.path.join(GOLD, "stats_zinc_v2_reference.npz"))
    graphs = [g for g in S.mol_graphs(z["atoms"], z["bonds"], S.ZINC_ATOMS) if len(g[0]) > 0][:8]
    ref = vectorize(mg._mol_nx(graphs, False), complexity=4, discrete=True).toarray()
    ours = S.nspdk_matrix(graphs).toarray()
    assert ref.shape == ours.shape and np.abs(ref - ours).max() < 1e-14 and (ref != 0).sum() == (ours != 0).sum()
