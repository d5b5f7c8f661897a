"""ActionBench Chamfer on the GPU (-m gpu): actionmesh_b200.evaluation against the reference's KD-tree implementation —
the golden values stored by the reference's own actionbench/chamfer.py (tests/golden/autoencoder_tiny.pt) and the oracle
restatement on larger clouds.  fp32 brute force vs fp64 KD-tree: relative difference <= 1e-5; nearest-neighbour indices equal
wherever the two nearest candidates are not tied to fp32 precision."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import autoencoder_oracle as ao

pytestmark = pytest.mark.gpu


def test_chamfer_score_matches_reference_goldens(amb_lib):
    from actionmesh_b200.evaluation import compute_chamfer_score

    g = load_golden("autoencoder_tiny.pt")
    a, b = g["chamfer_a"], g["chamfer_b"]
    assert abs(compute_chamfer_score(a, b, n=300) - g["chamfer_n300"]) <= 1e-5 * g["chamfer_n300"]
    assert abs(compute_chamfer_score(a, b, n=0) - g["chamfer_all"]) <= 1e-5 * g["chamfer_all"]


def test_nearest_neighbors_and_scores_on_large_clouds(amb_lib):
    from scipy.spatial import KDTree

    from actionmesh_b200 import ops
    from actionmesh_b200.evaluation import compute_chamfer_score, compute_motion_chamfer_score

    rng = np.random.default_rng(5)
    pred = rng.normal(size=(23_457, 3)).astype(np.float32)
    gt = (rng.normal(size=(100_003, 3)) * 1.1).astype(np.float32)
    d_ref, i_ref = KDTree(gt).query(pred)
    d, i = ops.nearest_neighbors(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda())
    assert np.allclose(d.cpu().numpy(), d_ref, rtol=1e-5, atol=1e-7)
    assert (i.cpu().numpy() == i_ref).mean() > 0.9999          # (fp32 ties aside)
    want = ao.chamfer_score(pred, gt, n=10_000, seed=44)
    assert abs(compute_chamfer_score(pred, gt, n=10_000, seed=44) - want) <= 1e-5 * want
    # single query / single reference / query count not a multiple of the block size
    d1, i1 = ops.nearest_neighbors(torch.from_numpy(pred[:1]).cuda(), torch.from_numpy(gt[:1]).cuda())
    assert int(i1) == 0 and abs(float(d1) - np.linalg.norm(pred[0] - gt[0])) < 1e-6
    # motion Chamfer: frame-0 correspondences, distances over the sequence (chamfer.py:56-89 restated with the KD-tree)
    T = 4
    P = (rng.normal(size=(T, 3000, 3))).astype(np.float32)
    G = (rng.normal(size=(T, 3500, 3))).astype(np.float32)
    _, ig = KDTree(P[0]).query(G[0])
    _, ip = KDTree(G[0]).query(P[0])
    ref = float(np.linalg.norm(P[:, ig] - G, axis=-1).mean(axis=0).mean() + np.linalg.norm(G[:, ip] - P, axis=-1).mean(axis=0).mean())
    assert abs(compute_motion_chamfer_score(P, G) - ref) <= 1e-5 * ref
