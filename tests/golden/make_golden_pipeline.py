"""Golden vectors for `PointSample` (the step before the hot path).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_pipeline.py
Output (committed): tests/golden/golden_point_sample.npz

mmdet3d's PointSample transform is third-party, but the reference carries a FIRST-PARTY copy of it
(projects/mmdet3d_plugin/models/detectors/uni3detr.py:50-111, class PointSample, used by the OV detector).
That class is executed here from the reference's own file (third-party imports stubbed by make_golden.py's
machinery) under a fixed legacy numpy seed, and its choices are stored. oracle/pipeline.py's restatement is
checked against them in tests/test_prestage.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG.install_stubs()
    det = MG.load_ref("projects/mmdet3d_plugin/models/detectors/uni3detr.py", "ref_detector_ps")
    out = {}
    for ci, (n, num, seed) in enumerate([(5000, 2000, 0), (300, 1000, 1), (1000, 1000, 2), (7, 3, 3)]):
        pts = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
        ps = det.PointSample(num_points=num)
        np.random.seed(seed)
        sampled, choices = ps._points_random_sampling(pts, num, ps.sample_range, ps.replace, return_choices=True)
        out[f"c{ci}_n"], out[f"c{ci}_num"], out[f"c{ci}_seed"] = n, num, seed
        out[f"c{ci}_choices"] = np.asarray(choices)
        assert np.array_equal(sampled, pts[choices])
        print(f"case {ci}: {n} -> {num}, replace={n < num}, unique={len(np.unique(choices))}")
    np.savez_compressed(os.path.join(HERE, "golden_point_sample.npz"), **out)


if __name__ == "__main__":
    main()
