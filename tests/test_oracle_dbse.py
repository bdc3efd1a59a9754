"""The oracle's restatement of DirectBSEdit (reference src/dbse.cpp), pinned on the CPU:
Eigen's HouseholderQR is absent, so the packed-QR conventions are pinned against LAPACK (scipy.linalg.qr mode="raw"),
the weight formula against its numpy statement, the displacement loop against an FP32 numpy loop."""
import numpy as np
import pytest
import scipy.linalg


def _case(seed, P, S):
    rng = np.random.default_rng(seed)
    rest = rng.standard_normal((P, 3)).astype(np.float32)
    shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3))).astype(np.float32)
    pos = (rest + 0.05 * rng.standard_normal((P, 3))).astype(np.float32)
    return rest, shapes, pos


@pytest.mark.parametrize("P,S", [(40, 1), (200, 7), (333, 24), (3, 5)])
def test_packed_qr_matches_lapack(oracle, P, S):
    rest, shapes, _ = _case(1, P, S)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    assert M.shape == (3 * P, S)
    # dbse.cpp:24-27: the delta is subtracted in FP32 and widened
    np.testing.assert_array_equal(M, (shapes - rest[None]).reshape(S, -1).T.astype(np.float64))
    QR, tau = oracle.householder_qr(M)
    (qr_l, tau_l), _ = scipy.linalg.qr(M, mode="raw")
    k = min(M.shape)
    np.testing.assert_allclose(QR, qr_l, rtol=0, atol=1e-12 * np.abs(qr_l).max())
    np.testing.assert_allclose(tau[:k], tau_l[:k], rtol=0, atol=1e-13)
    # the packed factors reproduce M: Q R = M
    Q, R = scipy.linalg.qr(M, mode="economic")
    np.testing.assert_allclose(np.abs(np.triu(QR[:k])), np.abs(R[:k]), atol=1e-12 * np.abs(R).max())


def test_zero_tail_column_gives_identity_reflector(oracle):
    M = np.zeros((9, 2), order="F")
    M[0, 0] = 2.0          # column 0 has a zero tail: tau = 0, beta = alpha (Eigen / dgeqr2)
    M[:, 1] = np.arange(9)
    QR, tau = oracle.householder_qr(M)
    assert tau[0] == 0.0 and QR[0, 0] == 2.0
    (qr_l, tau_l), _ = scipy.linalg.qr(M, mode="raw")
    np.testing.assert_allclose(QR, qr_l, atol=1e-13)


def test_weights_and_displace_follow_the_reference_loops(oracle):
    P, S = 150, 6
    rest, shapes, pos = _case(2, P, S)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR, _ = oracle.householder_qr(M)
    w = oracle.dbse_weights(QR, pos, rest)
    delta = (pos - rest).astype(np.float64).reshape(-1)           # :46-48
    np.testing.assert_allclose(w, delta @ QR, rtol=1e-12, atol=1e-14)  # (delta.asDiagonal() * matrixQR()).colwise().sum()
    for wr, dofall, fr in ((None, 0, 1.0), ((0.0, 1.0), 1, 0.5), ((-0.2, 0.3), 1, 0.0)):
        out = oracle.dbse_displace(M, w, pos, rest, weightrange=wr, dofalloff=dofall, falloffradius=fr)
        ref = np.zeros((P, 3), dtype=np.float32)
        for s in range(S):                                        # dbse.cpp:63-73, FP32, column order
            cw = np.float32(w[s] * 3)
            if wr is not None:
                cw = np.float32(min(max(cw, np.float32(wr[0])), np.float32(wr[1])))
            ref = (ref + (M[:, s].astype(np.float32).reshape(P, 3) * cw).astype(np.float32)).astype(np.float32)
        if dofall and fr != 0.0:                                  # SOP_FaceDeform.cpp:467-470
            ref = (ref + ((pos - rest) * np.float32(fr)).astype(np.float32)).astype(np.float32)
        np.testing.assert_array_equal(out, (rest + ref).astype(np.float32))


def test_identity_when_pose_is_rest(oracle):
    rest, shapes, _ = _case(3, 60, 4)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR, _ = oracle.householder_qr(M)
    w = oracle.dbse_weights(QR, rest, rest)
    assert np.all(w == 0.0)
    np.testing.assert_array_equal(oracle.dbse_displace(M, w, rest, rest), rest)


def test_packed_qr_reconstructs_the_shapes_matrix(oracle):
    """Q R = M with Q = H_0 ... H_{S-1}, H_j = I - tau_j v_j v_j^T read back from the packed storage (v_j has an implied
    1 on the diagonal and the stored essential part below it): the packed matrix means what dbse.cpp:53 assumes."""
    rest, shapes, _ = _case(9, 120, 6)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR, tau = oracle.householder_qr(M)
    m, n = M.shape
    X = np.triu(QR[:n]).copy()                       # R, then apply H_{n-1} ... H_0 from the left
    X = np.vstack([X, np.zeros((m - n, n))])
    for j in range(n - 1, -1, -1):
        v = np.zeros(m)
        v[j] = 1.0
        v[j + 1:] = QR[j + 1:, j]
        X -= tau[j] * np.outer(v, v @ X)
    np.testing.assert_allclose(X, M, rtol=0, atol=1e-12 * np.abs(M).max())


def test_weights_are_linear_in_the_pose_delta(oracle):
    rest, shapes, pos = _case(10, 80, 5)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR, _ = oracle.householder_qr(M)
    d = (pos - rest)
    w1 = oracle.dbse_weights(QR, rest + d, rest)
    w2 = oracle.dbse_weights(QR, rest + 2 * d, rest)   # 2 * d is exact in FP32, rest + 2 d may round: compare loosely
    np.testing.assert_allclose(w2, 2 * w1, rtol=0, atol=2e-6 * np.abs(w1).max())
