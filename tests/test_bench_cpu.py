"""Host-side pieces of bench.py / benchmarks that need no GPU: the weak-scaling workload (the single-GPU stroke list repeated
per band), the reference arm's own stroke expansion, and the batched stroke expansion of the C ABI."""
import numpy as np


def test_weak_scaling_workload_repeats_the_single_gpu_list(built_lib):
    import bench

    _, rec1, cx1, cy1, th1, radii1 = bench.build_workload(40)
    strokes, rec, cx, cy, th, radii = bench.build_workload(40, tiles=3)
    assert len(rec) == 3 * len(rec1) and radii == radii1
    for i in (0, 7, 39):
        a1, m1 = int(rec1["first_imprint"][i]), int(rec1["n_imprints"][i])
        for k in range(3):
            j = 3 * i + k
            a, m = int(rec["first_imprint"][j]), int(rec["n_imprints"][j])
            assert abs(m - m1) <= 2  # int(|p1 - p0|) may tip over where the shifted coordinates round differently
            assert rec["radius"][j] == rec1["radius"][i] and np.array_equal(rec["K"][j], rec1["K"][i])
            n = min(m, m1, 20)
            assert np.allclose(cx[a:a + n], cx1[a1:a1 + n], atol=1e-6) and np.allclose(cy[a:a + n], cy1[a1:a1 + n] + k * bench.ROWS, atol=1e-6)


def test_reference_arm_expansion_equals_the_library(built_lib, port):
    import bench
    from painty_b200 import api

    strokes = bench.build_strokes(12)
    for s in strokes[::3]:
        want = api.expand_stroke(s["path"], mode=0)
        got = bench.expand_with_oracle(port, s["path"])
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    assert int(bench.imprint_counts(strokes).sum()) == sum(len(api.expand_stroke(s["path"])[0]) for s in strokes)


def test_batched_stroke_expansion(built_lib):
    from painty_b200 import api

    r = np.random.default_rng(0)
    paths = [r.uniform(0, 500, (int(r.integers(1, 9)), 2)) for _ in range(40)]  # a 1-vertex stroke has no imprints
    fv = np.cumsum([0] + [len(p) for p in paths[:-1]])
    cx, cy, th, fi, ni = api.expand_strokes(fv, [len(p) for p in paths], np.concatenate(paths))
    assert fi[0] == 0 and fi[-1] + ni[-1] == len(cx)
    for i, p in enumerate(paths):
        a, b, c = api.expand_stroke(p)
        assert len(a) == ni[i]
        assert np.array_equal(a, cx[fi[i]:fi[i] + ni[i]]) and np.array_equal(b, cy[fi[i]:fi[i] + ni[i]]) and np.array_equal(c, th[fi[i]:fi[i] + ni[i]])
