"""Evaluation metrics: the reference's alignment classes (utils.py:540-801)
and eval.py's compute_v2v.  The numpy restatement (oracle/metrics_port.py) is pinned against
outputs of the unmodified reference classes (tests/golden/ref_metrics.npz); the CUDA kernel
(sfx_aligned_errors) is checked against the same vectors."""
import numpy as np
import pytest

from oracle import metrics_port as MP
from tests import common as Cm


def test_port_reproduces_the_reference_classes():
    g = Cm.golden('ref_metrics.npz')
    est, gt, vids = g['est'], g['gt'], g['vids']
    for b in range(est.shape[0]):
        assert np.allclose(MP.procrustes(est[b], gt[b]), g['procrustes_aligned'][b], rtol=0, atol=1e-12)
        assert np.allclose(MP.scale_align(est[b], gt[b]), g['scale_aligned'][b], rtol=0, atol=1e-13)
        assert np.array_equal(MP.point_error(est[b], gt[b]), g['none'][b])
        assert np.allclose(MP.pelvis_error(est[b], gt[b]), g['pelvis'][b], rtol=0, atol=1e-14)
    v = MP.compute_v2v(est, gt)
    assert np.allclose(v['procrustes'], g['procrustes'], rtol=0, atol=1e-12)
    v = MP.compute_v2v(est, gt, vids)
    assert np.allclose(v['procrustes'], g['procrustes_vids'], rtol=0, atol=1e-12)
    assert np.allclose(v['pelvis'], g['pelvis_vids'], rtol=0, atol=1e-14)
    # frame 1 is a reflection: no rotation aligns it, the others align to the noise level
    assert g['procrustes'][1].mean() > 10 * g['procrustes'][0].mean()
    assert g['procrustes'][0].mean() < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize('dt,tol', [('float64', 1e-11), ('float32', 2e-5)])
def test_device_metrics_against_the_reference(dt, tol):
    import torch
    from smplifyx_b200 import metrics as M, _native as N
    g = Cm.golden('ref_metrics.npz')
    est, gt, vids = g['est'].astype(dt), g['gt'].astype(dt), g['vids']
    pa, pe = M.ProcrustesAlignmentMPJPE(), M.PelvisAlignmentMPJPE()
    assert np.abs(pa(est, gt)['point'].cpu().numpy() - g['procrustes']).max() < tol
    assert np.abs(pe(est, gt)['point'].cpu().numpy() - g['pelvis']).max() < tol
    assert np.abs(M.mpjpe(est, gt).cpu().numpy() - g['none']).max() < tol
    assert np.abs(M.ProcrustesAlignment()(est, gt).cpu().numpy() - g['procrustes_aligned']).max() < tol
    assert np.abs(M.ScaleAlignment()(est, gt).cpu().numpy() - g['scale_aligned']).max() < tol
    # one frame, the reference's call shape ([N,3])
    one = pa(est[2], gt[2])['point']
    assert one.shape == (est.shape[1],) and np.abs(one.cpu().numpy() - g['procrustes'][2]).max() < tol
    out = M.compute_v2v(torch.tensor(est), torch.tensor(gt), {'procrustes': pa, 'pelvis': pe}, vids=vids)
    assert np.abs(out['point']['procrustes'] - g['procrustes_vids']).max() < tol
    assert np.abs(out['point']['pelvis'] - g['pelvis_vids']).max() < tol
    assert set(out.keys()) == {'point', 'fscore'}
    with pytest.raises(NotImplementedError):
        M.ProcrustesAlignmentMPJPE(fscore_thresholds=[0.005])(est, gt)


@pytest.mark.gpu
def test_device_metrics_on_fitted_meshes():
    """PA-V2V of a full SMPL-X-sized mesh batch (10 475 vertices): invariant under a similarity
    transform of the estimate, zero for identical sets."""
    import torch
    from smplifyx_b200 import metrics as M
    rng = np.random.default_rng(3)
    gt = torch.tensor(rng.normal(size=(16, 10475, 3)), dtype=torch.float32, device='cuda')
    R = torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]], device='cuda')
    est = 1.7 * gt @ R.T + torch.tensor([0.3, -2.0, 5.0], device='cuda')
    pa = M.ProcrustesAlignmentMPJPE()
    assert float(pa(est, gt)['point'].max()) < 2e-5
    noisy = est + 0.01 * torch.tensor(rng.normal(size=est.shape), dtype=torch.float32, device='cuda')
    e = pa(noisy, gt)['point']
    assert e.shape == (16, 10475) and 0.005 < float(e.mean()) / 1.7 < 0.02
