"""CPU tests: the oracle (the reference's own serial ELAS compiled into oracle/_ref, strict IEEE) is pinned by the
committed golden fixtures that tests/golden/make_golden.py produced from the reference in the build container.

The reference's own tests hold no known-answer vector for this path (SURVEY.md finding 6); the known answers the
survey recorded from the compiled reference are asserted here as well."""
import hashlib

import numpy as np
import pytest

import parity


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def preset(ref, name):
    return ref.params(0) if name == "robotics" else ref.pipeline_params()


def test_oracle_build_flags_are_strict_ieee(ref, golden_meta):
    # SURVEY.md finding 2: the parity oracle must not be a -ffast-math / -march=native build
    assert "-ffast-math" not in ref.flags and "-march" not in ref.flags
    assert "-ffp-contract=off" in ref.flags
    assert ref.flags == golden_meta["oracle_flags"]


@pytest.mark.parametrize("pname", ["robotics", "pipeline"])
def test_oracle_stage_taps_match_golden_hashes(ref, kitti_gray, golden, golden_meta, pname):
    """Every stage tap of the oracle, re-run here, hashes to what the fixture generator recorded."""
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    t = ref.staged(preset(ref, pname), L, R)
    want = golden_meta["hashes"]["%s_0" % pname]
    assert set(want) <= set(t)
    bad = [k for k in want if sha(t[k]) != want[k]]
    assert not bad, "oracle taps differ from the committed golden hashes: %s" % bad
    assert np.array_equal(t["support"], golden["%s_0_support" % pname])
    assert np.array_equal(t["tri1"], golden["%s_0_tri1" % pname])
    assert np.array_equal(t["tri2"], golden["%s_0_tri2" % pname])
    assert np.array_equal(t["D1"], golden["%s_0_D1" % pname])
    assert np.array_equal(t["D1raw"].astype(np.int16), golden["%s_0_D1raw" % pname])


def test_oracle_process_equals_staged_and_is_deterministic(ref, kitti_gray, golden):
    L, R = kitti_gray["L7"], kitti_gray["R7"]
    p = ref.pipeline_params()
    D1a, D2a, _ = ref.process(p, L, R)
    D1b, D2b, _ = ref.process(p, L, R)
    assert np.array_equal(D1a, D1b) and np.array_equal(D2a, D2b)
    assert np.array_equal(D1a, golden["pipeline_7_D1"])


def test_survey_known_answers(golden):
    """numbers recorded from the compiled reference on kitti_mini frame 0."""
    assert len(golden["robotics_0_support"]) == 1254
    assert len(golden["robotics_0_tri1"]) == 2469 and len(golden["robotics_0_tri2"]) == 2469
    assert int(golden["robotics_0_support"][:, 2].max()) == 88
    assert int((golden["robotics_0_D1"] >= 0).sum()) == 315374
    assert len(golden["pipeline_0_support"]) == 1952
    assert len(golden["pipeline_0_tri1"]) == 3896
    assert int((golden["pipeline_0_D1"] >= 0).sum()) == 465750


def test_oracle_reproduces_dataset_digests(ref, kitti_gray):
    """The oracle on ALL of the reference's own inputs -- the 21 pairs of datasets/kitti_mini (driver preset) and the 7 pairs of
    datasets/profile (runProfiling's parameters, stereo_vision.cu:727-730) -- against the digests that
    tests/golden/make_dataset_fixtures.py recorded from it (support points, triangles, valid pixels, sha256 of both maps)."""
    import json
    import os

    from conftest import GOLDEN

    dig = json.load(open(os.path.join(GOLDEN, "dataset_digests.json")))
    assert ref.flags == dig["oracle_flags"]
    for i in (0, 5, 13, 20):  # a spread of the kitti pairs here; the -m gpu tests run all 21 against the live oracle
        D1, D2, _ = ref.process(ref.pipeline_params(), kitti_gray["L%d" % i], kitti_gray["R%d" % i])
        want = dig["kitti_pipeline"][str(i)]
        assert (sha(D1), sha(D2), int((D1 >= 0).sum())) == (want["D1"], want["D2"], want["valid1"]), "kitti pair %d" % i
    z = np.load(os.path.join(GOLDEN, "profile_gray.npz"))
    assert sorted(dig["profile_runprofiling"]) == sorted({k[:-2] for k in z.files})
    for name in ("cones", "urban3"):
        L, R = z[name + "_L"], z[name + "_R"]
        assert list(L.shape) == dig["profile_runprofiling"][name]["shape"]
        D1, D2, _ = ref.process(ref.params(0, postprocess_only_left=0), L, R)
        want = dig["profile_runprofiling"][name]
        assert (sha(D1), sha(D2), int((D1 >= 0).sum()), int((D2 >= 0).sum())) == (want["D1"], want["D2"], want["valid1"], want["valid2"]), name


def test_oracle_prior_table(ref):
    """P[] of elas.cpp:831 as listed in SURVEY.md 8a row 11, restated in numpy."""
    for setting, want in ((0, [-14, -9, -2, 0]), (1, [-9, -5, -1, 0])):
        p = ref.params(setting)
        tss = 2 * p.sigma * p.sigma
        got = [int((-np.log(p.gamma + np.exp(-d * d / tss)) + np.log(p.gamma)) / p.beta) for d in range(4)]
        assert got == want


def test_reproject_restatement_known_point(golden_meta):
    """The CPU restatement of stereo_vision.cu:188-212,324 (u8 conversion in numpy, projection in oracle/project_port.c) on hand-computed points."""
    Q = np.array(golden_meta["Q"])
    assert abs(Q[0, 3] + 738.7995529175) < 1e-6 and abs(Q[2, 3] - 1027.855158176) < 1e-6 and abs(Q[3, 2] - 1.861616069957) < 1e-9
    D = np.full((2, 3), -10.0, np.float32)
    D[0, 0] = 10.0  # -> u8 40
    D[1, 2] = 0.125  # 0.5 -> round half even -> 0  -> w = 0 -> inf
    D[0, 1] = 0.375  # 1.5 -> 2
    D[1, 1] = 70.0  # saturates at 255
    d8, pts = parity.reproject_oracle(D, Q, np.eye(3), np.zeros(3))
    assert d8.tolist() == [[40, 2, 0], [0, 255, 0]]
    w = Q[3, 2] * 40
    assert np.allclose(pts[0], [Q[0, 3] / w, Q[1, 3] / w, Q[2, 3] / w], rtol=1e-15)
    assert not np.isfinite(pts[5]).all()
    XR, XT = np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
    _, pts2 = parity.reproject_oracle(D, Q, XR, XT)
    assert np.allclose(pts2[0], XR.reshape(3, 3) @ pts[0] + XT.reshape(3), rtol=1e-13)


def test_project_port_against_uncontracted_formula(golden, golden_meta):
    """oracle/project_port.c (projectParallel with the reference build's fused multiply-adds) against the formula as written in the
    source, every operation rounded on its own (numpy): identical inf / NaN pattern, differences of rounding size only; and the
    port's fma is a true single-rounding one (a product whose low half decides the result)."""
    from oracle.ref import port_project

    Q = np.array(golden_meta["Q"])
    Qg = Q.copy()
    Qg[3, 0], Qg[3, 1], Qg[3, 3] = 1e-4, -3e-4, 0.37
    XR, XT = np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
    for D, q in ((golden["pipeline_0_D1"], Q), (golden["robotics_7_D1"], Qg)):
        d8, got = parity.reproject_oracle(D, q, XR, XT)
        want = parity.reproject_numpy(d8.astype(np.float64), q, XR, XT)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        # scale of the terms that enter the sums: cancellation near the principal point makes the RELATIVE difference of a
        # coordinate large-ish (1e-13), the difference against the size of the summands is a few ulps
        scale = np.abs(want[fin]).max()
        assert np.abs(got[fin] - want[fin]).max() <= 64 * np.finfo(np.float64).eps * scale
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)
        assert rel.max() <= 1e-9
    # single rounding: a = 1 + 2^-30, b = 1 - 2^-30, a * b = 1 - 2^-60 exactly: fma(a, b, -1) = -2^-60 where two roundings give 0.
    # point[0] = fma(XR02, Z, fma(XR00, X, XR01 * Y)) + XT0 with XR00 = a, X = b, XR01 = 1, Y = -1
    a, b = 1.0 + 2.0**-30, 1.0 - 2.0**-30
    Qs = np.zeros((4, 4))
    Qs[0, 3], Qs[1, 3], Qs[3, 3] = b, -1.0, 1.0
    XRs = np.array([[a, 1.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    pts = port_project(np.zeros((1, 1)), 1, 1, Qs, XRs, np.zeros(3))
    assert pts[0, 0] == -(2.0**-60) and a * b - 1.0 == 0.0
    # pos[0] = fma(d, Q02, fma(x, Q00, y * Q01)) + Q03 with d = a, Q02 = b, y * Q01 = -1 (y = 1)
    Qs = np.zeros((4, 4))
    Qs[0, 2], Qs[0, 1], Qs[3, 3] = b, -1.0, 1.0
    pts = port_project(np.array([[0.0], [a]]), 2, 1, Qs, np.eye(3), np.zeros(3))
    assert pts[1, 0] == -(2.0**-60)


def test_project_port_against_exact_rational_arithmetic(golden, golden_meta):
    """oracle/project_port.c against the same operation sequence evaluated in exact rational arithmetic with one rounding where the
    reference build has one (DMUL / DFMA / DADD / division = round-to-nearest-even of the exact result; float(Fraction) rounds
    correctly): independent of libm's fma and of the compiler, on 600 pixels of a real map with a general Q."""
    from fractions import Fraction as F

    from oracle.ref import port_project

    Q = np.array(golden_meta["Q"])
    Q[3, 0], Q[3, 1], Q[3, 3] = 1e-4, -3e-4, 0.37
    XR, XT = np.array(golden_meta["XR"]).reshape(3, 3), np.array(golden_meta["XT"]).reshape(3)
    D = golden["robotics_7_D1"]
    H, W = D.shape
    d8, pts = parity.reproject_oracle(D, Q, XR, XT)
    assert np.array_equal(pts, port_project(d8.astype(np.float64), H, W, Q, XR, XT), equal_nan=True)

    def rn(x):
        return float(x)

    def fma(a, b, c):
        return rn(F(a) * F(b) + F(c))

    rng = np.random.default_rng(17)
    idx = rng.choice(np.flatnonzero(d8.reshape(-1) > 0), 600, replace=False)
    for p in idx:
        y, x = divmod(int(p), W)
        d = float(d8.reshape(-1)[p])
        pos = [rn(F(fma(d, Q[j, 2], fma(float(x), Q[j, 0], rn(F(float(y)) * F(Q[j, 1]))))) + F(Q[j, 3])) for j in range(4)]
        assert pos[3] != 0.0
        X, Y, Z = (rn(F(pos[k]) / F(pos[3])) for k in range(3))
        want = [rn(F(fma(XR[j, 2], Z, fma(XR[j, 0], X, rn(F(XR[j, 1]) * F(Y))))) + F(XT[j])) for j in range(3)]
        assert pts[p].tolist() == want, (p, pts[p], want)
