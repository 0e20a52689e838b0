"""Generates tests/golden/bench_parity.npz: the BASELINE bench workloads (bench.py) fitted by the
oracle port of the reference (oracle/fit_port.py, pinned to the unmodified reference by
tests/test_oracle.py) on a fixed sample of frames, together with the oracle-vs-oracle envelope
of the same frames (the same fit started from keypoints one float32 ulp apart), so that
tests/test_gpu_bench_parity.py can hold the CUDA engine to the reference's own run-to-run spread
on exactly the workloads the benchmark times.

    python tests/golden/make_bench_parity.py          # ~10 minutes on 8 cores

Workloads (bench.bench_cfg): cfg2 (default line: no regression prior, guess_init), cfg2_reg
(combined regression + camera prior), cfg3 (VPoser, 5 stages), cfg5 (8 frames of rank 3's shard
of the 8 x 512 regression-prior job).  Keypoints come from the oracle's own forward pass and are
stored in the fixture, so the GPU test needs neither the oracle nor /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench          # noqa: E402

VSTRIDE = 5           # every 5th vertex is kept


class _Args(object):
    pass


CASES = {
    'cfg2': dict(n=16, seed=0, kw=dict()),
    'cfg2_reg': dict(n=16, seed=0, kw=dict(regression_prior=True)),
    'cfg3': dict(n=8, seed=0, kw=dict(vposer=True)),
    'cfg5': dict(n=8, seed=3000, kw=dict(regression_prior=True)),
}


def one_ulp_apart(kp, direction=np.inf):
    out = kp.copy()
    out[:, :, :2] = np.nextafter(kp[:, :, :2], np.float32(direction))
    return out


def variants(kp):
    """The oracle's own runs of the same frames: the inputs themselves, one float32 ulp up, one
    ulp down, and a relative 1e-6 apart."""
    rel = kp.copy()
    rel[:, :, :2] *= np.float32(1.0 + 1e-6)
    return [('ref', kp), ('ref_ulp', one_ulp_apart(kp)), ('ref_ulp2', one_ulp_apart(kp, -np.inf)),
            ('ref_ulp3', rel)]


def main(only=None):
    only = [o for o in (only or []) if o != 'redo']
    path = os.path.join(HERE, 'bench_parity.npz')
    out = dict(np.load(path, allow_pickle=False)) if os.path.isfile(path) else {}
    procs = os.cpu_count() or 1
    for name, case in CASES.items():
        if only and name not in only:
            continue
        cfg = bench.bench_cfg(**case['kw'])
        a = _Args()
        a.frames, a.seed = case['n'], case['seed']
        kp, expose, pixie = bench.make_inputs_cpu(a, cfg)
        frames = list(range(case['n']))
        runs, tags = [], []
        for tag, k in variants(kp):
            tags.append(tag)
            if '{}/{}/loss'.format(name, tag) in out and 'redo' not in sys.argv:
                runs.append(None)
                continue
            pool = bench.OraclePool(cfg, k, expose, pixie, min(procs, len(frames)))
            _, res = pool.round(frames, want_result=True)
            pool.close()
            runs.append(res)
        out[name + '/keypoints'] = kp.astype(np.float32)
        if expose is not None:
            for key in ('body_pose', 'global_orient', 'transl', 'center'):
                out[name + '/expose/' + key] = np.stack([np.asarray(e[key]) for e in expose])
            for key in ('body_pose', 'global_pose'):
                out[name + '/pixie/' + key] = np.stack([np.asarray(p[key]) for p in pixie])
        for tag, res in zip(tags, runs):
            if res is None:
                continue
            out['{}/{}/loss'.format(name, tag)] = np.array([r['loss'] for r in res])
            out['{}/{}/vertices'.format(name, tag)] = np.stack(
                [r['vertices'][::VSTRIDE] for r in res]).astype(np.float32)
            out['{}/{}/joints'.format(name, tag)] = np.stack([r['joints'] for r in res])
            out['{}/{}/cam_t'.format(name, tag)] = np.stack([r['cam_t'] for r in res])
            out['{}/{}/center'.format(name, tag)] = np.stack([r['center'] for r in res])
            out['{}/{}/evals'.format(name, tag)] = np.array([r['evals'] for r in res])
            out['{}/{}/n_orient'.format(name, tag)] = np.array([r['n_orient'] for r in res])
        print(name, [int(out['{}/{}/evals'.format(name, t)].mean()) for t in tags])
        np.savez_compressed(path, **out)


if __name__ == '__main__':
    main(sys.argv[1:])
