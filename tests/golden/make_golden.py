"""Regenerates tests/golden/reference_fixtures.json.

The reference (pure Julia) cannot run in this image, and its tests compute their expectations at run time with
serial Julia (`y_ref = A * x_global`), they do not store literals.  Every fixture below therefore records the
LITERAL INPUTS of one reference test (file:line cited) and an expectation computed here with DENSE numpy linear
algebra — the same mathematical object serial Julia computes, obtained without any code under oracle/ or the
product.  The VectorPlan / ghost-map tables are the worked examples of SURVEY.md App. C (derived there from the
reference's src/sparse.jl:1875-1984 by a separate throw-away restatement); no reference test reads those fields.

Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def dense(I, J, V, m, n, dtype):
    A = np.zeros((m, n), dtype=dtype)
    for i, j, v in zip(I, J, V):
        A[i - 1, j - 1] += v  # Julia sparse(I,J,V) sums duplicates
    return A


def enc(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return {"re": a.real.tolist(), "im": a.imag.tolist()}
    return a.tolist()


def tridiagonal(n, cplx):
    """test/test_utils.jl:90-100 tridiagonal_matrix(T, n)"""
    I = list(range(1, n + 1)) + list(range(1, n)) + list(range(2, n + 1))
    J = list(range(1, n + 1)) + list(range(2, n + 1)) + list(range(1, n))
    V = [2.0] * n + [-0.5] * (n - 1) + [-0.5] * (n - 1)
    if cplx:
        Vi = [0.1] * n + [0.2] * (n - 1) + [-0.2] * (n - 1)
        V = [complex(a, b) for a, b in zip(V, Vi)]
    return I, J, V


def test_vector(n, cplx):
    """test/test_utils.jl:124-130 test_vector(T, n)"""
    if cplx:
        return [complex(k, n - k + 1) for k in range(1, n + 1)]
    return [float(k) for k in range(1, n + 1)]


def main():
    fx = []
    for cplx in (False, True):
        T = "c128" if cplx else "f64"
        dt = np.complex128 if cplx else np.float64

        # --- test/test_vector_multiplication.jl:41-65 (A*x) and :70-92 (mul!) ------------------------------------
        I, J, V = tridiagonal(8, cplx)
        x = np.array(test_vector(8, cplx), dtype=dt)
        A = dense(I, J, V, 8, 8, dt)
        case = {
            "name": f"tridiag8_{T}",
            "source": "test/test_vector_multiplication.jl:41-65,70-92; test/test_utils.jl:90-100,124-130",
            "dtype": T, "m": 8, "n": 8, "nranks": 2, "I": I, "J": J, "V": enc(V), "x": enc(x),
            "y": enc(A @ x),
            "yT": enc(A.T @ x),  # transpose(x)*A, test_vector_multiplication.jl:141-150 (complex only there)
            "tol": 1e-10,  # test/test_utils.jl:154-157
        }
        if cplx:
            case["y_adj"] = enc(A.T @ np.conj(x))  # x'*A, test_vector_multiplication.jl:152-159
        fx.append(case)

        # --- test/test_vector_multiplication.jl:97-118 non-square 6x8 ---------------------------------------------
        I = [1, 2, 3, 4, 5, 6, 1, 2, 3, 4]
        J = [1, 2, 3, 4, 5, 6, 7, 8, 1, 2]
        V = [complex(k, 11 - k) for k in range(1, 11)] if cplx else [float(k) for k in range(1, 11)]
        x = np.array(test_vector(8, cplx), dtype=dt)
        A = dense(I, J, V, 6, 8, dt)
        fx.append({
            "name": f"nonsquare6x8_{T}", "source": "test/test_vector_multiplication.jl:97-118",
            "dtype": T, "m": 6, "n": 8, "nranks": 2, "I": I, "J": J, "V": enc(V), "x": enc(x),
            "y": enc(A @ x), "tol": 1e-10,
        })

        # --- test/test_local_constructors.jl:214-229 10x8 with empty rows -----------------------------------------
        I = [1, 2, 3, 4, 5, 6, 1, 3, 5, 7, 9]
        J = [1, 2, 3, 4, 5, 6, 6, 5, 4, 3, 2]
        V = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.5, 0.5, 0.5, 0.5, 0.5]
        x = np.arange(1, 9).astype(dt)
        A = dense(I, J, V, 10, 8, dt)
        fx.append({
            "name": f"local10x8_{T}", "source": "test/test_local_constructors.jl:214-229",
            "dtype": T, "m": 10, "n": 8, "nranks": 2, "I": I, "J": J, "V": enc(np.array(V, dtype=dt)), "x": enc(x),
            "y": enc(A @ x), "tol": 1e-10,
        })

        # --- test/test_repartition.jl:122-126,178-187 8x6, non-uniform row partition [1,6,9] at 2 ranks ------------
        I = [1, 2, 3, 4, 5, 6, 7, 8, 1, 3, 5, 7, 2, 4, 6, 8]
        J = [1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6, 4, 5, 6, 1]
        V = [float(k) for k in range(1, 17)]
        x = np.ones(6, dtype=dt)
        A = dense(I, J, V, 8, 6, dt)
        fx.append({
            "name": f"repart8x6_{T}", "source": "test/test_repartition.jl:122-151,178-187",
            "dtype": T, "m": 8, "n": 6, "nranks": 2, "I": I, "J": J, "V": enc(np.array(V, dtype=dt)), "x": enc(x),
            "row_partition": [1, 6, 9], "y": enc(A @ x), "tol": 1e-10,
        })

        # --- test/test_new_operations.jl:46-54,66-76,139-147 symmetric 8x8, transpose(A)*x, dot ---------------------
        I = list(range(1, 9)) * 2
        J = list(range(1, 9)) + [2, 3, 4, 5, 6, 7, 8, 1]
        V = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
        S = dense(I, J, V, 8, 8, dt)
        A = S + S.T + 2 * np.eye(8, dtype=dt)
        Id, Jd = np.nonzero(A)
        x = (np.arange(1, 9) + 0.1).astype(dt)
        y0 = (np.arange(8, 0, -1) + 0.1).astype(dt)
        fx.append({
            "name": f"sym8_{T}", "source": "test/test_new_operations.jl:46-54,66-76,139-147",
            "dtype": T, "m": 8, "n": 8, "nranks": 2, "I": (Id + 1).tolist(), "J": (Jd + 1).tolist(),
            "V": enc(A[Id, Jd]), "x": enc(x), "y": enc(A @ x), "yT": enc(A.T @ x),
            "dot_with": enc(y0), "dot_xy": enc(np.vdot(x, y0)), "dot_xx": enc(np.vdot(x, x)), "tol": 1e-10,
        })

        # --- test/test_transpose.jl:39-54 10x8 TransposePlan ---------------------------------------------------------
        I = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 1, 3, 5, 7, 9]
        J = [1, 2, 3, 4, 5, 6, 7, 8, 1, 2, 3, 5, 7, 1, 4]
        V = [complex(k, 16 - k) for k in range(1, 16)] if cplx else [float(k) for k in range(1, 16)]
        A = dense(I, J, V, 10, 8, dt)
        x10 = np.array(test_vector(10, cplx), dtype=dt)
        fx.append({
            "name": f"transpose10x8_{T}", "source": "test/test_transpose.jl:39-54",
            "dtype": T, "m": 10, "n": 8, "nranks": 2, "I": I, "J": J, "V": enc(V),
            "x": enc(np.array(test_vector(8, cplx), dtype=dt)), "y": enc(A @ np.array(test_vector(8, cplx), dtype=dt)),
            "xT": enc(x10), "yT": enc(A.T @ x10), "AT_dense": enc(A.T), "tol": 1e-10,
        })

        # --- test/test_transpose.jl:59-79 square 8x8, diag 2.0, upper 0.3, lower 0.7 ------------------------------------
        n2 = 8
        I = list(range(1, n2 + 1)) + list(range(1, n2)) + list(range(2, n2 + 1))
        J = list(range(1, n2 + 1)) + list(range(2, n2 + 1)) + list(range(1, n2))
        V = [2.0] * n2 + [0.3] * (n2 - 1) + [0.7] * (n2 - 1)
        if cplx:
            Vi = [0.1] * n2 + [-0.1] * (n2 - 1) + [0.2] * (n2 - 1)
            V = [complex(a, b) for a, b in zip(V, Vi)]
        A = dense(I, J, V, n2, n2, dt)
        x = np.array(test_vector(8, cplx), dtype=dt)
        fx.append({
            "name": f"nonsym8_{T}", "source": "test/test_transpose.jl:59-79",
            "dtype": T, "m": 8, "n": 8, "nranks": 2, "I": I, "J": J, "V": enc(V), "x": enc(x),
            "y": enc(A @ x), "xT": enc(x), "yT": enc(A.T @ x), "AT_dense": enc(A.T), "tol": 1e-10,
        })

    # Worked plan tables, SURVEY.md App. C (C.1, C.3, C.5, C.7) — 2 ranks, 1-based.
    plans = {
        "tridiag8": {  # C.1 (structure identical for f64 and c128)
            "rank0": {"rowptr": [1, 3, 6, 9, 12], "colval": [1, 2, 1, 2, 3, 2, 3, 4, 3, 4, 5], "col_indices": [1, 2, 3, 4, 5],
                      "recv_rank_ids": [1], "recv_perm": [[5]], "send_rank_ids": [1], "send_indices": [[4]],
                      "local_src": [1, 2, 3, 4], "local_dst": [1, 2, 3, 4]},
            "rank1": {"rowptr": [1, 4, 7, 10, 12], "colval": [1, 2, 3, 2, 3, 4, 3, 4, 5, 4, 5], "col_indices": [4, 5, 6, 7, 8],
                      "recv_rank_ids": [0], "recv_perm": [[1]], "send_rank_ids": [0], "send_indices": [[1]],
                      "local_src": [1, 2, 3, 4], "local_dst": [2, 3, 4, 5]},
        },
        "nonsquare6x8": {  # C.3
            "rank0": {"rowptr": [1, 3, 5, 7], "colval": [1, 4, 2, 5, 1, 3], "col_indices": [1, 2, 3, 7, 8],
                      "recv_rank_ids": [1], "recv_perm": [[4, 5]], "send_rank_ids": [1], "send_indices": [[2, 4]],
                      "local_src": [1, 2, 3], "local_dst": [1, 2, 3]},
            "rank1": {"rowptr": [1, 3, 4, 5], "colval": [1, 2, 3, 4], "col_indices": [2, 4, 5, 6],
                      "recv_rank_ids": [0], "recv_perm": [[1, 2]], "send_rank_ids": [0], "send_indices": [[3, 4]],
                      "local_src": [1, 2], "local_dst": [3, 4]},
        },
        "repart8x6": {  # C.5
            "rank0": {"rowptr": [1, 3, 5, 7, 9, 10], "colval": [1, 3, 2, 4, 3, 4, 4, 5, 5], "col_indices": [1, 2, 3, 4, 5],
                      "recv_rank_ids": [1], "recv_perm": [[4, 5]], "send_rank_ids": [1], "send_indices": [[1, 2]],
                      "local_src": [1, 2, 3], "local_dst": [1, 2, 3]},
            "rank1": {"rowptr": [1, 2, 4, 6], "colval": [3, 1, 3, 1, 2], "col_indices": [1, 2, 6],
                      "recv_rank_ids": [0], "recv_perm": [[1, 2]], "send_rank_ids": [0], "send_indices": [[1, 2]],
                      "local_src": [3], "local_dst": [3]},
        },
        "transpose10x8_AT": {  # C.7: the materialised transpose (real values V=1:15)
            "rank0": {"rowptr": [1, 4, 6, 8, 10], "colval": [1, 5, 6, 2, 7, 1, 3, 4, 6], "col_indices": [1, 2, 3, 4, 7, 9, 10],
                      "nzval": [1, 14, 9, 2, 10, 11, 3, 4, 15]},
            "rank1": {"rowptr": [1, 3, 4, 6, 7], "colval": [1, 2, 3, 2, 4, 5], "col_indices": [3, 5, 6, 7, 8],
                      "nzval": [12, 5, 6, 13, 7, 8]},
        },
    }
    # ---- the callers of the hot path: literal inputs of the reference's product tests ------------------
    products = []
    for cplx in (False, True):
        T = "c128" if cplx else "f64"
        dt = np.complex128 if cplx else np.float64
        n = 8
        # test/test_matrix_multiplication.jl:38-62: tridiagonal A (test_utils.jl:90-100) times a second tridiagonal B
        IA, JA, VA = tridiagonal(n, cplx)
        IB = list(range(1, n + 1)) + list(range(1, n)) + list(range(2, n + 1))
        JB = list(range(1, n + 1)) + list(range(2, n + 1)) + list(range(1, n))
        VB = [1.5] * n + [0.25] * (n - 1) + [0.25] * (n - 1)
        if cplx:
            VB = [complex(a, b) for a, b in zip(VB, [-0.1] * n + [0.1] * (n - 1) + [0.1] * (n - 1))]
        A, B = dense(IA, JA, VA, n, n, dt), dense(IB, JB, VB, n, n, dt)
        products.append({"name": f"spgemm_tridiag8_{T}", "kind": "sparse*sparse", "source": "test/test_matrix_multiplication.jl:38-62", "dtype": T,
                         "A": {"m": n, "n": n, "I": IA, "J": JA, "V": enc(VA)}, "B": {"m": n, "n": n, "I": IB, "J": JB, "V": enc(VB)},
                         "C": enc(A @ B), "tol": 1e-10})
        # test/test_matrix_multiplication.jl:65-92: 6x8 times 8x10
        IA2, JA2 = [1, 2, 3, 4, 5, 6, 1, 2, 3, 4], [1, 2, 3, 4, 5, 6, 7, 8, 1, 2]
        IB2, JB2 = [1, 2, 3, 4, 5, 6, 7, 8, 1, 3], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
        V2 = [complex(k, 11 - k) for k in range(1, 11)] if cplx else [float(k) for k in range(1, 11)]
        A2, B2 = dense(IA2, JA2, V2, 6, 8, dt), dense(IB2, JB2, V2, 8, 10, dt)
        products.append({"name": f"spgemm_6x8x10_{T}", "kind": "sparse*sparse", "source": "test/test_matrix_multiplication.jl:65-92", "dtype": T,
                         "A": {"m": 6, "n": 8, "I": IA2, "J": JA2, "V": enc(V2)}, "B": {"m": 8, "n": 10, "I": IB2, "J": JB2, "V": enc(V2)},
                         "C": enc(A2 @ B2), "tol": 1e-10})
        # test/test_new_operations.jl:46-59, 79-88: A_sparse = S + S^T + 2I times the dense 8x6 B; also transpose(A) * B
        IS = [1, 2, 3, 4, 5, 6, 7, 8, 1, 2, 3, 4, 5, 6, 7, 8]
        JS = [1, 2, 3, 4, 5, 6, 7, 8, 2, 3, 4, 5, 6, 7, 8, 1]
        VS = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
        Sd = dense(IS, JS, VS, n, n, dt)
        Ad = Sd + Sd.T + 2.0 * np.eye(n, dtype=dt)
        I3, J3 = np.nonzero(Ad)
        Bd = np.array([[i + j * 0.1 for j in range(1, 7)] for i in range(1, 9)], dtype=dt)
        products.append({"name": f"spmm_sym8_{T}", "kind": "sparse*dense", "source": "test/test_new_operations.jl:46-59, 79-88", "dtype": T,
                         "A": {"m": n, "n": n, "I": (I3 + 1).tolist(), "J": (J3 + 1).tolist(), "V": enc(Ad[I3, J3])}, "Bdense": enc(Bd),
                         "C": enc(Ad @ Bd), "CT": enc(Ad.T @ Bd), "tol": 1e-10})
    with open(os.path.join(HERE, "reference_fixtures.json"), "w") as f:
        json.dump({"fixtures": fx, "plans": plans, "products": products}, f, indent=1)
    print(f"wrote {len(fx)} fixtures, {len(products)} product fixtures")


if __name__ == "__main__":
    main()
