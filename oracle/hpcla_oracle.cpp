// oracle/hpcla_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the distributed sparse mat-vec hot path of HPCLinearAlgebra.jl
// (sloisel/LinearAlgebraMPI.jl).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the CUDA product under
// linearalgebrampi.jl_b200/ never does.
//
// Parity status: the reference is pure Julia and cannot run in this image (no julia, no MPI).
// The arithmetic is pinned against the reference's own test fixtures (tests/golden/, taken from
// test/test_vector_multiplication.jl, test_transpose.jl, test_local_constructors.jl,
// test_repartition.jl, test_new_operations.jl).  The PLAN ARRAYS (VectorPlan fields, ghost maps)
// are read by no reference test: for them parity is UNPINNED by the reference and rests on this
// line-by-line restatement, its independent numpy twin (oracle/oracle.py) and the structural
// invariants of the algorithm.
//
// Every function cites the reference file:line it follows.  All index arrays are 1-based and
// int64 at this API (the reference's Ti-typed arrays hold the same values); the SpMV kernels are
// additionally instantiated for int32 so the CPU baseline streams the real index width.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off: the reference's `acc += a*x` is not
// contracted to an FMA by Julia, so neither is this).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

typedef int64_t i64;

// ------------------------------------------------------------------------------------------------
// src/HPCLinearAlgebra.jl:279-289  uniform_partition(n, nranks)
// ------------------------------------------------------------------------------------------------
extern "C" void orc_uniform_partition(i64 n, i64 nranks, i64* partition) {
    i64 per_rank = n / nranks;
    i64 remainder = n % nranks;
    partition[0] = 1;
    for (i64 r = 1; r <= nranks; ++r) {
        i64 extra = (r <= remainder) ? 1 : 0;
        partition[r] = partition[r - 1] + per_rank + extra;
    }
}

// Julia's searchsortedlast(v, x): index (1-based) of the last element <= x, 0 if none.
static inline i64 searchsortedlast(const i64* v, i64 len, i64 x) {
    return (i64)(std::upper_bound(v, v + len, x) - v);
}
// Julia's searchsortedfirst(v, x): index (1-based) of the first element >= x, len+1 if none.
static inline i64 searchsortedfirst(const i64* v, i64 len, i64 x) {
    return (i64)(std::lower_bound(v, v + len, x) - v) + 1;
}

// src/sparse.jl:1890-1894  owner of global index g in a partition (with the clamp)
extern "C" i64 orc_owner(const i64* partition, i64 nranks, i64 g) {
    i64 owner = searchsortedlast(partition, nranks + 1, g) - 1;
    if (owner >= nranks) owner = nranks - 1;
    return owner;
}

// ------------------------------------------------------------------------------------------------
// src/sparse.jl:501 and :137-144   col_indices = unique!(sort(copy(rowval)));
//                                  colval[k] = searchsortedfirst(col_indices, rowval[k])
// global_cols: 1-based global column of every stored entry (row-major order).
// col_indices_out has capacity nnz; returns ncols_compressed.
// ------------------------------------------------------------------------------------------------
extern "C" i64 orc_compress(i64 nnz, const i64* global_cols, i64* col_indices_out, i64* colval_out) {
    if (nnz == 0) return 0;
    std::vector<i64> s(global_cols, global_cols + nnz);
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    i64 ncc = (i64)s.size();
    std::copy(s.begin(), s.end(), col_indices_out);
    for (i64 k = 0; k < nnz; ++k) colval_out[k] = searchsortedfirst(s.data(), ncc, global_cols[k]);
    return ncc;
}

// ------------------------------------------------------------------------------------------------
// src/sparse.jl:1875-1984  VectorPlan(A, x) — built for all P ranks at once; the Alltoall of
// counts (:1899-1900) and the tag-20 index messages (:1908-1936) become in-process moves.
// ------------------------------------------------------------------------------------------------
struct RankPlan {
    std::vector<i64> send_rank_ids;
    std::vector<std::vector<i64>> send_indices;
    std::vector<i64> recv_rank_ids;
    std::vector<std::vector<i64>> recv_perm;
    std::vector<i64> local_src, local_dst;
    i64 n_gathered = 0;
};
struct OrcPlans {
    i64 P = 0;
    std::vector<RankPlan> ranks;
};

extern "C" OrcPlans* orc_plans_build(i64 P, const i64* const* col_indices, const i64* ncc, const i64* x_partition) {
    OrcPlans* W = new OrcPlans();
    W->P = P;
    W->ranks.resize(P);
    // needed_from[r][o] = (global_idx, dst_idx) pairs rank r needs from owner o   (:1886-1895)
    std::vector<std::vector<std::vector<std::pair<i64, i64>>>> needed(P);
    for (i64 r = 0; r < P; ++r) {
        needed[r].resize(P);
        W->ranks[r].n_gathered = ncc[r];
        for (i64 d = 0; d < ncc[r]; ++d) {
            i64 g = col_indices[r][d];
            i64 owner = orc_owner(x_partition, P, g);
            needed[r][owner].push_back({g, d + 1});
        }
    }
    for (i64 rank = 0; rank < P; ++rank) {
        RankPlan& pl = W->ranks[rank];
        i64 my_x_start = x_partition[rank];
        // Step 3 (:1908-1918): who do I receive from, and where does it land in `gathered`
        for (i64 r = 0; r < P; ++r) {
            if (!needed[rank][r].empty() && r != rank) {
                pl.recv_rank_ids.push_back(r);
                std::vector<i64> dst;
                for (auto& t : needed[rank][r]) dst.push_back(t.second);
                pl.recv_perm.push_back(dst);
            }
        }
        // Step 4/5 (:1925-1944): requests that reach me; global -> local index into x.v
        for (i64 r = 0; r < P; ++r) {
            if (!needed[r][rank].empty() && r != rank) {
                pl.send_rank_ids.push_back(r);
                std::vector<i64> loc;
                for (auto& t : needed[r][rank]) loc.push_back(t.first - my_x_start + 1);
                pl.send_indices.push_back(loc);
            }
        }
        // Step 6 (:1947-1953): elements I own
        for (auto& t : needed[rank][rank]) {
            pl.local_src.push_back(t.first - my_x_start + 1);
            pl.local_dst.push_back(t.second);
        }
        // Step 7 (:1956-1957) sort!(rank ids): already ascending by construction.
    }
    return W;
}
extern "C" void orc_plans_free(OrcPlans* W) { delete W; }

// field ids: 0 send_rank_ids, 1 recv_rank_ids, 2 local_src, 3 local_dst, 4 send_indices[slot], 5 recv_perm[slot]
static const std::vector<i64>* plan_field(const OrcPlans* W, i64 rank, int field, i64 slot) {
    const RankPlan& p = W->ranks[rank];
    switch (field) {
        case 0: return &p.send_rank_ids;
        case 1: return &p.recv_rank_ids;
        case 2: return &p.local_src;
        case 3: return &p.local_dst;
        case 4: return &p.send_indices[slot];
        case 5: return &p.recv_perm[slot];
    }
    return nullptr;
}
extern "C" i64 orc_plans_len(const OrcPlans* W, i64 rank, int field, i64 slot) {
    return (i64)plan_field(W, rank, field, slot)->size();
}
extern "C" void orc_plans_get(const OrcPlans* W, i64 rank, int field, i64 slot, i64* out) {
    const std::vector<i64>* v = plan_field(W, rank, field, slot);
    std::copy(v->begin(), v->end(), out);
}

// ------------------------------------------------------------------------------------------------
// src/vectors.jl:394-463  execute_plan!(plan, x) for all ranks; Isend/Irecv (tag 21) become
// memcpy between rank buffers.  Byte-generic over the element type.
// ------------------------------------------------------------------------------------------------
static void execute_rank(const OrcPlans* W, i64 rank, const char* const* x_local, char* const* gathered, i64 es,
                         std::vector<std::vector<char>>* sendbufs /* per (rank) list of peer buffers */) {
    const RankPlan& p = W->ranks[rank];
    // Step 1 (:426-428)
    for (size_t i = 0; i < p.local_src.size(); ++i)
        std::memcpy(gathered[rank] + (p.local_dst[i] - 1) * es, x_local[rank] + (p.local_src[i] - 1) * es, es);
    // Step 2 (:431-439): pack
    std::vector<std::vector<char>>& bufs = sendbufs[rank];
    bufs.resize(p.send_rank_ids.size());
    for (size_t i = 0; i < p.send_rank_ids.size(); ++i) {
        const std::vector<i64>& idx = p.send_indices[i];
        bufs[i].resize(idx.size() * es);
        for (size_t k = 0; k < idx.size(); ++k) std::memcpy(bufs[i].data() + k * es, x_local[rank] + (idx[k] - 1) * es, es);
    }
}
static void scatter_rank(const OrcPlans* W, i64 rank, char* const* gathered, i64 es,
                         const std::vector<std::vector<char>>* sendbufs) {
    const RankPlan& p = W->ranks[rank];
    // Steps 3-4 (:442-455): the message from peer r is r's send buffer addressed to me
    for (size_t i = 0; i < p.recv_rank_ids.size(); ++i) {
        i64 r = p.recv_rank_ids[i];
        const RankPlan& q = W->ranks[r];
        size_t slot = std::find(q.send_rank_ids.begin(), q.send_rank_ids.end(), rank) - q.send_rank_ids.begin();
        const std::vector<char>& buf = sendbufs[r][slot];
        const std::vector<i64>& perm = p.recv_perm[i];
        for (size_t k = 0; k < perm.size(); ++k) std::memcpy(gathered[rank] + (perm[k] - 1) * es, buf.data() + k * es, es);
    }
}
extern "C" void orc_plans_execute(const OrcPlans* W, const void* const* x_local, void* const* gathered, i64 elem_size) {
    std::vector<std::vector<std::vector<char>>> sendbufs(W->P);
    for (i64 r = 0; r < W->P; ++r) execute_rank(W, r, (const char* const*)x_local, (char* const*)gathered, elem_size, sendbufs.data());
    for (i64 r = 0; r < W->P; ++r) scatter_rank(W, r, (char* const*)gathered, elem_size, sendbufs.data());
}

// ------------------------------------------------------------------------------------------------
// src/sparse.jl:2055-2066  _spmv_kernel!: per row, acc = zero(T); acc += nzval[j]*x[colval[j]] over
// the stored entries left to right.  (= the arithmetic of mul!, src/sparse.jl:2019-2037.)
// Complex product is the textbook (ar*br - ai*bi) + i(ar*bi + ai*br), spelled out so that
// -ffp-contract=off and no library cmul semantics (inf/nan fix-ups) interfere.
// ------------------------------------------------------------------------------------------------
struct c128 { double re, im; };
static inline float mul_add(float acc, float a, float b) { float p = a * b; return acc + p; }
static inline double mul_add(double acc, double a, double b) { double p = a * b; return acc + p; }
static inline c128 mul_add(c128 acc, c128 a, c128 b) {
    double pr = a.re * b.re - a.im * b.im;
    double pi = a.re * b.im + a.im * b.re;
    return c128{acc.re + pr, acc.im + pi};
}
template <class T> static inline T zero_of();
template <> inline float zero_of<float>() { return 0.0f; }
template <> inline double zero_of<double>() { return 0.0; }
template <> inline c128 zero_of<c128>() { return c128{0.0, 0.0}; }

template <class T, class Ti>
static void spmv_rows(i64 row_begin, i64 row_end, const Ti* rowptr, const Ti* colval, const T* nzval, const T* x, T* y) {
    for (i64 row = row_begin; row < row_end; ++row) {
        T acc = zero_of<T>();
        for (i64 j = (i64)rowptr[row]; j <= (i64)rowptr[row + 1] - 1; ++j) acc = mul_add(acc, nzval[j - 1], x[(i64)colval[j - 1] - 1]);
        y[row] = acc;
    }
}
#define ORC_SPMV(NAME, T, Ti)                                                                                          \
    extern "C" void NAME(i64 nrows, const Ti* rowptr, const Ti* colval, const T* nzval, const T* gathered, T* y) {   \
        spmv_rows<T, Ti>(0, nrows, rowptr, colval, nzval, gathered, y);                                              \
    }
ORC_SPMV(orc_spmv_f32_i32, float, int32_t)
ORC_SPMV(orc_spmv_f32_i64, float, int64_t)
ORC_SPMV(orc_spmv_f64_i32, double, int32_t)
ORC_SPMV(orc_spmv_f64_i64, double, int64_t)
ORC_SPMV(orc_spmv_c128_i32, c128, int32_t)
ORC_SPMV(orc_spmv_c128_i64, c128, int64_t)

// ------------------------------------------------------------------------------------------------
// src/vectors.jl:798-812 dot (Julia's dot conjugates its FIRST argument) and :758-766 norm(v,2):
// local reduction then an Allreduce(+).  The reference's local reductions are BLAS (summation
// order unspecified) => compared with a relative tolerance, never bitwise.  Here: plain
// left-to-right sums, ranks added in rank order.
// ------------------------------------------------------------------------------------------------
extern "C" double orc_dot_f64(i64 n, const double* x, const double* y) {
    double s = 0;
    for (i64 i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
extern "C" float orc_dot_f32(i64 n, const float* x, const float* y) {
    float s = 0;
    for (i64 i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
extern "C" void orc_dot_c128(i64 n, const c128* x, const c128* y, double* out2) {
    double sr = 0, si = 0;
    for (i64 i = 0; i < n; ++i) {  // conj(x)*y
        sr += x[i].re * y[i].re + x[i].im * y[i].im;
        si += x[i].re * y[i].im - x[i].im * y[i].re;
    }
    out2[0] = sr;
    out2[1] = si;
}

// ------------------------------------------------------------------------------------------------
// src/sparse.jl:1551-1744 TransposePlan(A) and :1756-1829 execute_plan!(plan, A), for all ranks.
// Inputs per rank: 1-based rowptr / compressed colval / col_indices / nzval (byte-generic).
// Result per rank: CSR of the owned rows of A^T (row_partition = A.col_partition), ascending
// columns, compressed columns, values moved (never conjugated).
// ------------------------------------------------------------------------------------------------
struct RankT {
    std::vector<i64> rowptr, colval, col_indices, global_cols;
    std::vector<char> nzval;
    // value-movement plan (TransposePlan fields), exported for completeness
    std::vector<i64> rank_ids, recv_rank_ids, local_src, local_dst;
    std::vector<std::vector<i64>> send_indices, recv_perm;
};
struct OrcTranspose {
    i64 P = 0;
    std::vector<RankT> ranks;
};

extern "C" OrcTranspose* orc_transpose_build(i64 P, const i64* row_partition, const i64* col_partition, const i64* const* rowptr,
                                             const i64* const* colval, const i64* const* col_indices,
                                             const void* const* nzval, i64 es) {
    OrcTranspose* W = new OrcTranspose();
    W->P = P;
    W->ranks.resize(P);
    // Step 1 (:1569-1579): send_to[rank][dest] = (global_row, j, idx)
    typedef std::tuple<i64, i64, i64> T3;
    std::vector<std::vector<std::vector<T3>>> send_to(P);
    for (i64 rank = 0; rank < P; ++rank) {
        send_to[rank].resize(P);
        i64 my_row_start = row_partition[rank];
        i64 nrows_local = row_partition[rank + 1] - row_partition[rank];
        for (i64 local_col = 1; local_col <= nrows_local; ++local_col) {
            i64 global_row = my_row_start + local_col - 1;
            for (i64 idx = rowptr[rank][local_col - 1]; idx <= rowptr[rank][local_col] - 1; ++idx) {
                i64 local_j = colval[rank][idx - 1];
                i64 j = col_indices[rank][local_j - 1];
                i64 dest_rank = searchsortedlast(col_partition, P + 1, j) - 1;
                send_to[rank][dest_rank].push_back(T3(global_row, j, idx));
            }
        }
    }
    for (i64 rank = 0; rank < P; ++rank) {
        RankT& R = W->ranks[rank];
        // Step 3 (:1589-1606)
        for (i64 r = 0; r < P; ++r)
            if (!send_to[rank][r].empty() && r != rank) {
                R.rank_ids.push_back(r);
                std::vector<i64> ind;
                for (auto& t : send_to[rank][r]) ind.push_back(std::get<2>(t));
                R.send_indices.push_back(ind);
            }
        // Step 4 (:1613-1621)
        for (i64 r = 0; r < P; ++r)
            if (!send_to[r][rank].empty() && r != rank) R.recv_rank_ids.push_back(r);
        // Step 5 (:1626-1655): entries (j, i, source_rank, source_idx)
        i64 my_AT_row_start = col_partition[rank];
        i64 my_AT_row_end = col_partition[rank + 1] - 1;
        i64 local_ncols = my_AT_row_end - my_AT_row_start + 1;
        typedef std::tuple<i64, i64, i64, i64> T4;
        std::vector<T4> entries;
        for (i64 r : R.recv_rank_ids) {
            i64 k = 0;
            for (auto& t : send_to[r][rank]) entries.push_back(T4(std::get<1>(t), std::get<0>(t), r, ++k));
        }
        std::vector<i64> local_entries_src;
        for (auto& t : send_to[rank][rank]) {
            local_entries_src.push_back(std::get<2>(t));
            entries.push_back(T4(std::get<1>(t), std::get<0>(t), rank, (i64)local_entries_src.size()));
        }
        std::stable_sort(entries.begin(), entries.end(), [&](const T4& a, const T4& b) {
            i64 ja = std::get<0>(a) - my_AT_row_start + 1, jb = std::get<0>(b) - my_AT_row_start + 1;
            if (ja != jb) return ja < jb;
            return std::get<1>(a) < std::get<1>(b);
        });
        // CSC of result.AT == CSR of the owned rows of A^T (:1659-1688)
        std::vector<i64> colptr(local_ncols + 1, 0), rowval(entries.size());
        for (auto& e : entries) colptr[std::get<0>(e) - my_AT_row_start + 1] += 1;
        colptr[0] = 1;
        for (i64 c = 1; c <= local_ncols; ++c) colptr[c] += colptr[c - 1];
        std::vector<i64> cursors(colptr.begin(), colptr.end() - 1), entry_to_nz(entries.size());
        for (size_t e = 0; e < entries.size(); ++e) {
            i64 lc = std::get<0>(entries[e]) - my_AT_row_start + 1;
            i64 pos = cursors[lc - 1];
            rowval[pos - 1] = std::get<1>(entries[e]);
            entry_to_nz[e] = pos;
            cursors[lc - 1] += 1;
        }
        // Step 6 (:1693-1714): value plan
        R.recv_perm.resize(R.recv_rank_ids.size());
        for (size_t i = 0; i < R.recv_rank_ids.size(); ++i) R.recv_perm[i].assign(send_to[R.recv_rank_ids[i]][rank].size(), 0);
        for (size_t e = 0; e < entries.size(); ++e) {
            i64 src_rank = std::get<2>(entries[e]), src_idx = std::get<3>(entries[e]), dst = entry_to_nz[e];
            if (src_rank == rank) {
                R.local_src.push_back(local_entries_src[src_idx - 1]);
                R.local_dst.push_back(dst);
            } else {
                size_t slot = std::find(R.recv_rank_ids.begin(), R.recv_rank_ids.end(), src_rank) - R.recv_rank_ids.begin();
                R.recv_perm[slot][src_idx - 1] = dst;
            }
        }
        // col_indices of the result (:1723) and compression (:1802, compress_AT_cached)
        R.global_cols = rowval;
        std::vector<i64> ci(rowval);
        std::sort(ci.begin(), ci.end());
        ci.erase(std::unique(ci.begin(), ci.end()), ci.end());
        R.col_indices = ci;
        R.rowptr = colptr;
        R.colval.resize(rowval.size());
        for (size_t k = 0; k < rowval.size(); ++k) R.colval[k] = searchsortedfirst(ci.data(), (i64)ci.size(), rowval[k]);
        R.nzval.assign(entries.size() * es, 0);
    }
    // execute_plan! (:1756-1796): move the values
    for (i64 rank = 0; rank < P; ++rank) {
        RankT& R = W->ranks[rank];
        const char* src = (const char*)nzval[rank];
        for (size_t i = 0; i < R.local_src.size(); ++i)
            std::memcpy(R.nzval.data() + (R.local_dst[i] - 1) * es, src + (R.local_src[i] - 1) * es, es);
        for (size_t i = 0; i < R.rank_ids.size(); ++i) {
            i64 r = R.rank_ids[i];
            RankT& Q = W->ranks[r];
            size_t slot = std::find(Q.recv_rank_ids.begin(), Q.recv_rank_ids.end(), rank) - Q.recv_rank_ids.begin();
            const std::vector<i64>& perm = Q.recv_perm[slot];
            const std::vector<i64>& sidx = R.send_indices[i];
            for (size_t k = 0; k < sidx.size(); ++k)
                std::memcpy(Q.nzval.data() + (perm[k] - 1) * es, src + (sidx[k] - 1) * es, es);
        }
    }
    return W;
}
extern "C" void orc_transpose_free(OrcTranspose* W) { delete W; }
// field ids: 0 rowptr, 1 colval (compressed), 2 col_indices, 3 global_cols, 4 nzval (bytes)
extern "C" i64 orc_transpose_len(const OrcTranspose* W, i64 rank, int field) {
    const RankT& R = W->ranks[rank];
    switch (field) {
        case 0: return (i64)R.rowptr.size();
        case 1: return (i64)R.colval.size();
        case 2: return (i64)R.col_indices.size();
        case 3: return (i64)R.global_cols.size();
        case 4: return (i64)R.nzval.size();
    }
    return -1;
}
extern "C" void orc_transpose_get(const OrcTranspose* W, i64 rank, int field, void* out) {
    const RankT& R = W->ranks[rank];
    switch (field) {
        case 0: std::memcpy(out, R.rowptr.data(), R.rowptr.size() * 8); break;
        case 1: std::memcpy(out, R.colval.data(), R.colval.size() * 8); break;
        case 2: std::memcpy(out, R.col_indices.data(), R.col_indices.size() * 8); break;
        case 3: std::memcpy(out, R.global_cols.data(), R.global_cols.size() * 8); break;
        case 4: std::memcpy(out, R.nzval.data(), R.nzval.size()); break;
    }
}

// ------------------------------------------------------------------------------------------------
// CPU baseline: the reference's distributed A*x on host cores.  One worker thread stands for one
// MPI rank (one contiguous row block).  Per repetition each worker runs the execute_plan! sequence
// of src/vectors.jl:394-463 — local copy into gathered_cpu, pack, "Isend/Irecv" (a copy out of the
// peer's packed buffer after a barrier), scatter, and the second full copy gathered_cpu -> gathered
// that vectors.jl:460 performs on the CPU path — then the row-serial loop of src/sparse.jl:2055-2066.
// Protocol mirrors tools/benchmark_vs_petsc.jl:64-75: barrier, timed repetition, barrier.
// times_out[rep] = wall seconds of repetition rep (max over workers by construction of the barriers).
// ------------------------------------------------------------------------------------------------
struct SpinBarrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit SpinBarrier(int n_) : n(n_) {}
    void wait() {
        int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            count.store(0, std::memory_order_relaxed);
            gen.fetch_add(1, std::memory_order_release);
        } else {
            int spins = 0;
            while (gen.load(std::memory_order_acquire) == g)
                if (++spins > 2000) std::this_thread::yield();
        }
    }
};

template <class T, class Ti>
static void bench_worker(const OrcPlans* W, i64 rank, i64 nrows, const Ti* rowptr, const Ti* colval, const T* nzval, const T* x,
                         T* y, std::vector<std::vector<std::vector<T>>>* sendbufs, std::vector<std::vector<T>>* gathered_cpu,
                         std::vector<std::vector<T>>* gathered, SpinBarrier* bar, i64 warmup, i64 reps, double* times_out) {
    const RankPlan& p = W->ranks[rank];
    std::vector<T>& gc = (*gathered_cpu)[rank];
    std::vector<T>& g = (*gathered)[rank];
    std::vector<std::vector<T>>& mine = (*sendbufs)[rank];
    for (i64 it = 0; it < warmup + reps; ++it) {
        bar->wait();
        auto t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < p.local_src.size(); ++i) gc[p.local_dst[i] - 1] = x[p.local_src[i] - 1];
        for (size_t i = 0; i < p.send_rank_ids.size(); ++i) {
            const std::vector<i64>& idx = p.send_indices[i];
            T* buf = mine[i].data();
            for (size_t k = 0; k < idx.size(); ++k) buf[k] = x[idx[k] - 1];
        }
        bar->wait();  // stands for the completion of Isend/Irecv (Waitall, vectors.jl:446)
        for (size_t i = 0; i < p.recv_rank_ids.size(); ++i) {
            i64 r = p.recv_rank_ids[i];
            const RankPlan& q = W->ranks[r];
            size_t slot = std::find(q.send_rank_ids.begin(), q.send_rank_ids.end(), rank) - q.send_rank_ids.begin();
            const T* buf = (*sendbufs)[r][slot].data();
            const std::vector<i64>& perm = p.recv_perm[i];
            for (size_t k = 0; k < perm.size(); ++k) gc[perm[k] - 1] = buf[k];
        }
        std::memcpy((void*)g.data(), (const void*)gc.data(), gc.size() * sizeof(T));  // vectors.jl:460 via :163
        spmv_rows<T, Ti>(0, nrows, rowptr, colval, nzval, g.data(), y);
        bar->wait();
        auto t1 = std::chrono::steady_clock::now();
        if (rank == 0 && it >= warmup) times_out[it - warmup] = std::chrono::duration<double>(t1 - t0).count();
    }
}

template <class T, class Ti>
static void bench_run(const OrcPlans* W, const i64* nrows, const void* const* rowptr, const void* const* colval,
                      const void* const* nzval, const void* const* x, void* const* y, i64 warmup, i64 reps, double* times_out) {
    i64 P = W->P;
    std::vector<std::vector<std::vector<T>>> sendbufs(P);
    std::vector<std::vector<T>> gathered_cpu(P), gathered(P);
    for (i64 r = 0; r < P; ++r) {
        const RankPlan& p = W->ranks[r];
        sendbufs[r].resize(p.send_rank_ids.size());
        for (size_t i = 0; i < p.send_rank_ids.size(); ++i) sendbufs[r][i].resize(p.send_indices[i].size());
        gathered_cpu[r].resize(p.n_gathered);
        gathered[r].resize(p.n_gathered);
    }
    SpinBarrier bar((int)P);
    std::vector<std::thread> th;
    for (i64 r = 0; r < P; ++r)
        th.emplace_back(bench_worker<T, Ti>, W, r, nrows[r], (const Ti*)rowptr[r], (const Ti*)colval[r], (const T*)nzval[r],
                        (const T*)x[r], (T*)y[r], &sendbufs, &gathered_cpu, &gathered, &bar, warmup, reps, times_out);
    for (auto& t : th) t.join();
}

// dtype: 0 f32, 1 f64, 2 c128 ; itype: 0 int32, 1 int64.  Index arrays here are in the REAL width.
extern "C" int orc_bench_spmv(const OrcPlans* W, int dtype, int itype, const i64* nrows, const void* const* rowptr,
                              const void* const* colval, const void* const* nzval, const void* const* x, void* const* y,
                              i64 warmup, i64 reps, double* times_out) {
#define ORC_CASE(D, I, T, Ti) \
    if (dtype == D && itype == I) { bench_run<T, Ti>(W, nrows, rowptr, colval, nzval, x, y, warmup, reps, times_out); return 0; }
    ORC_CASE(0, 0, float, int32_t)
    ORC_CASE(0, 1, float, int64_t)
    ORC_CASE(1, 0, double, int32_t)
    ORC_CASE(1, 1, double, int64_t)
    ORC_CASE(2, 0, c128, int32_t)
    ORC_CASE(2, 1, c128, int64_t)
    return 1;
}

// ------------------------------------------------------------------------------------------------
// CPU reference arm on a SYNTHETIC workload, generated inside the workers.
// Same per-repetition work as bench_worker above, but every worker (= one reference MPI rank)
// allocates, generates and first-touches its own row block, vectors and buffers from its own
// thread (optionally pinned to one core), the way P MPI processes would: on a multi-socket host the
// pages then live next to the core that streams them.  The generators are the ones of
// include/hpcla_synth.h, handed in as function pointers (this library does not link the other).
//   kind 0..2: stencils on an nx x ny x nz grid;  kind 3: power-law rows (n = nx, seed, max_len)
//   op 0: `inner` multiplies per repetition (inner = ncols models the reference's column loop for
//         A * B::HPCMatrix, src/sparse.jl:2398-2403);
//   op 1: one CG iteration per repetition (q = A*p; alpha = rr / dot(p,q); x += alpha p;
//         r -= alpha q; rr' = dot(r,r); p = r + (rr'/rr) p — dot = local sum + Allreduce,
//         src/vectors.jl:798-812), real types only.
// times_out[rep]: wall seconds; nnz_out: global stored entries; ynorm2_out: sum |y|^2 of the last
// multiply (a cross-check value).
// ------------------------------------------------------------------------------------------------
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <cstdlib>

typedef i64 (*fn_stencil_nnz)(int, i64, i64, i64, i64, i64);
typedef int (*fn_stencil_fill)(int, i64, i64, i64, int, int, i64, i64, void*, void*, void*);
typedef i64 (*fn_powerlaw_nnz)(i64, uint64_t, i64, i64, i64);
typedef int (*fn_powerlaw_fill)(i64, uint64_t, i64, int, int, i64, i64, void*, void*, void*);
typedef int (*fn_vector)(int, uint64_t, i64, i64, void*);
typedef void (*fn_set_threads)(int);

struct SynthJob {
    fn_stencil_nnz stencil_nnz;
    fn_stencil_fill stencil_fill;
    fn_powerlaw_nnz powerlaw_nnz;
    fn_powerlaw_fill powerlaw_fill;
    fn_vector vector;
    fn_set_threads set_threads;
    int kind, dtype, itype, op, pin;
    i64 nx, ny, nz, n, P, warmup, reps, inner, max_len;
    uint64_t seed, x_seed;
    std::vector<i64> partition;
    // shared between the workers
    std::vector<std::vector<i64>> col_indices;
    std::vector<const i64*> ci_ptr;
    std::vector<i64> ncc;
    OrcPlans* plans = nullptr;
    std::vector<i64> nnz;
    std::vector<double> ynorm2, red;  // per-worker partial sums (Allreduce stand-in)
    std::atomic<int> failed{0};
    double* times_out;
};

static inline double norm2_of(float v) { return (double)v * (double)v; }
static inline double norm2_of(double v) { return v * v; }
static inline double norm2_of(c128 v) { return v.re * v.re + v.im * v.im; }

template <class T> struct Axpy {
    static void run(T*, const T*, double, i64) {}
    static double dot(const T*, const T*, i64) { return 0.0; }
    static void xpby(T*, const T*, double, i64) {}
};
template <> struct Axpy<float> {
    static void run(float* y, const float* x, double a, i64 n) { for (i64 i = 0; i < n; ++i) y[i] = y[i] + (float)a * x[i]; }
    static double dot(const float* x, const float* y, i64 n) { float s = 0; for (i64 i = 0; i < n; ++i) s += x[i] * y[i]; return (double)s; }
    static void xpby(float* p, const float* r, double b, i64 n) { for (i64 i = 0; i < n; ++i) p[i] = r[i] + (float)b * p[i]; }
};
template <> struct Axpy<double> {
    static void run(double* y, const double* x, double a, i64 n) { for (i64 i = 0; i < n; ++i) y[i] = y[i] + a * x[i]; }
    static double dot(const double* x, const double* y, i64 n) { double s = 0; for (i64 i = 0; i < n; ++i) s += x[i] * y[i]; return s; }
    static void xpby(double* p, const double* r, double b, i64 n) { for (i64 i = 0; i < n; ++i) p[i] = r[i] + b * p[i]; }
};

template <class T, class Ti>
static void synth_worker(SynthJob* J, i64 rank, SpinBarrier* bar, std::vector<std::vector<std::vector<T>>>* sendbufs) {
    if (J->pin) {
        const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET((int)(rank % (ncpu > 0 ? ncpu : 1)), &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);  // best effort
    }
    J->set_threads(1);
    const i64 rb = J->partition[rank] - 1, re = J->partition[rank + 1] - 1, nrows = re - rb;
    const i64 nnz = J->kind == 3 ? J->powerlaw_nnz(J->n, J->seed, J->max_len, rb, re) : J->stencil_nnz(J->kind, J->nx, J->ny, J->nz, rb, re);
    J->nnz[rank] = nnz;
    Ti* rowptr = (Ti*)std::malloc(sizeof(Ti) * (size_t)(nrows + 1));
    Ti* cols = (Ti*)std::malloc(sizeof(Ti) * (size_t)std::max<i64>(nnz, 1));
    T* vals = (T*)std::malloc(sizeof(T) * (size_t)std::max<i64>(nnz, 1));
    int rc = J->kind == 3 ? J->powerlaw_fill(J->n, J->seed, J->max_len, J->dtype, J->itype, rb, re, rowptr, cols, vals)
                          : J->stencil_fill(J->kind, J->nx, J->ny, J->nz, J->dtype, J->itype, rb, re, rowptr, cols, vals);
    if (rc || !rowptr || !cols || !vals) J->failed.store(1);
    // src/sparse.jl:501, 137-144: col_indices = unique!(sort(copy(cols))); colval = searchsortedfirst(col_indices, col)
    {
        std::vector<i64>& ci = J->col_indices[rank];
        ci.assign(cols, cols + nnz);
        std::sort(ci.begin(), ci.end());
        ci.erase(std::unique(ci.begin(), ci.end()), ci.end());
        ci.shrink_to_fit();
        const i64 ncc = (i64)ci.size();
        for (i64 k = 0; k < nnz; ++k) cols[k] = (Ti)searchsortedfirst(ci.data(), ncc, (i64)cols[k]);
        J->ncc[rank] = ncc;
        J->ci_ptr[rank] = ci.data();
    }
    bar->wait();
    if (rank == 0) J->plans = orc_plans_build(J->P, J->ci_ptr.data(), J->ncc.data(), J->partition.data());  // src/sparse.jl:1875-1984
    bar->wait();
    const RankPlan& p = J->plans->ranks[rank];
    std::vector<std::vector<T>>& mine = (*sendbufs)[rank];
    mine.resize(p.send_rank_ids.size());
    for (size_t i = 0; i < mine.size(); ++i) mine[i].assign(p.send_indices[i].size(), zero_of<T>());
    std::vector<T> gc((size_t)p.n_gathered, zero_of<T>()), g((size_t)p.n_gathered, zero_of<T>());
    std::vector<T> x((size_t)nrows), y((size_t)nrows, zero_of<T>());
    J->vector(J->dtype, J->x_seed, rb, re, x.data());
    std::vector<T> cg_x, cg_r;  // CG state (op 1): p = x (the multiplied vector), q = y
    double rr = 0.0;
    auto allreduce = [&](double local) {  // Allreduce(+): everybody publishes, everybody adds in rank order
        J->red[rank] = local;
        bar->wait();
        double s = 0.0;
        for (i64 r = 0; r < J->P; ++r) s += J->red[r];
        bar->wait();
        return s;
    };
    if (J->op == 1) {
        cg_x.assign((size_t)nrows, zero_of<T>());
        cg_r = x;
        rr = allreduce(Axpy<T>::dot(cg_r.data(), cg_r.data(), nrows));
    }
    auto multiply = [&]() {  // execute_plan! (src/vectors.jl:394-463) + the row loop (src/sparse.jl:2055-2066)
        for (size_t i = 0; i < p.local_src.size(); ++i) gc[p.local_dst[i] - 1] = x[p.local_src[i] - 1];
        for (size_t i = 0; i < p.send_rank_ids.size(); ++i) {
            const std::vector<i64>& idx = p.send_indices[i];
            T* buf = mine[i].data();
            for (size_t k = 0; k < idx.size(); ++k) buf[k] = x[idx[k] - 1];
        }
        bar->wait();  // stands for the completion of Isend/Irecv (Waitall, vectors.jl:446)
        for (size_t i = 0; i < p.recv_rank_ids.size(); ++i) {
            const i64 r = p.recv_rank_ids[i];
            const RankPlan& q = J->plans->ranks[r];
            const size_t slot = std::find(q.send_rank_ids.begin(), q.send_rank_ids.end(), rank) - q.send_rank_ids.begin();
            const T* buf = (*sendbufs)[r][slot].data();
            const std::vector<i64>& perm = p.recv_perm[i];
            for (size_t k = 0; k < perm.size(); ++k) gc[perm[k] - 1] = buf[k];
        }
        bar->wait();  // nobody repacks while a peer still reads its buffer
        std::memcpy((void*)g.data(), (const void*)gc.data(), gc.size() * sizeof(T));  // vectors.jl:460 via :163
        spmv_rows<T, Ti>(0, nrows, rowptr, cols, vals, g.data(), y.data());
    };
    for (i64 it = 0; it < J->warmup + J->reps; ++it) {
        bar->wait();
        auto t0 = std::chrono::steady_clock::now();
        if (J->op == 0) {
            for (i64 k = 0; k < J->inner; ++k) multiply();
        } else {
            multiply();  // q = A p
            const double pq = allreduce(Axpy<T>::dot(x.data(), y.data(), nrows));
            const double alpha = rr / pq;
            Axpy<T>::run(cg_x.data(), x.data(), alpha, nrows);
            Axpy<T>::run(cg_r.data(), y.data(), -alpha, nrows);
            const double rr_new = allreduce(Axpy<T>::dot(cg_r.data(), cg_r.data(), nrows));
            Axpy<T>::xpby(x.data(), cg_r.data(), rr_new / rr, nrows);
            rr = rr_new;
        }
        bar->wait();
        auto t1 = std::chrono::steady_clock::now();
        if (rank == 0 && it >= J->warmup) J->times_out[it - J->warmup] = std::chrono::duration<double>(t1 - t0).count();
    }
    double s = 0.0;
    for (i64 i = 0; i < nrows; ++i) s += norm2_of(y[(size_t)i]);
    J->ynorm2[rank] = s;
    bar->wait();
    std::free(rowptr);
    std::free(cols);
    std::free(vals);
}

template <class T, class Ti>
static int synth_run(SynthJob* J) {
    std::vector<std::vector<std::vector<T>>> sendbufs((size_t)J->P);
    SpinBarrier bar((int)J->P);
    std::vector<std::thread> th;
    for (i64 r = 0; r < J->P; ++r) th.emplace_back(synth_worker<T, Ti>, J, r, &bar, &sendbufs);
    for (auto& t : th) t.join();
    return J->failed.load();
}

extern "C" int orc_bench_synth(const void* const* fns6, int kind, i64 nx, i64 ny, i64 nz, uint64_t seed, i64 max_len, uint64_t x_seed, int dtype,
                               int itype, i64 P, int pin, int op, i64 inner, i64 warmup, i64 reps, double* times_out, i64* nnz_out,
                               double* ynorm2_out) {
    SynthJob J;
    J.stencil_nnz = (fn_stencil_nnz)fns6[0];
    J.stencil_fill = (fn_stencil_fill)fns6[1];
    J.powerlaw_nnz = (fn_powerlaw_nnz)fns6[2];
    J.powerlaw_fill = (fn_powerlaw_fill)fns6[3];
    J.vector = (fn_vector)fns6[4];
    J.set_threads = (fn_set_threads)fns6[5];
    J.kind = kind, J.dtype = dtype, J.itype = itype, J.op = op, J.pin = pin;
    J.nx = nx, J.ny = ny, J.nz = nz, J.P = P, J.warmup = warmup, J.reps = reps, J.inner = inner < 1 ? 1 : inner, J.max_len = max_len;
    J.seed = seed, J.x_seed = x_seed;
    J.n = kind == 3 ? nx : (kind == 0 ? nx * ny : nx * ny * nz);
    if (P < 1 || J.n < P || reps < 1 || (op == 1 && dtype == 2)) return 2;
    J.partition.resize((size_t)P + 1);
    orc_uniform_partition(J.n, P, J.partition.data());
    J.col_indices.resize((size_t)P);
    J.ci_ptr.assign((size_t)P, nullptr);
    J.ncc.assign((size_t)P, 0);
    J.nnz.assign((size_t)P, 0);
    J.ynorm2.assign((size_t)P, 0.0);
    J.red.assign((size_t)P, 0.0);
    J.times_out = times_out;
    int rc = 1;
#undef ORC_CASE
#define ORC_CASE(D, I, T, Ti) \
    if (dtype == D && itype == I) rc = synth_run<T, Ti>(&J);
    ORC_CASE(0, 0, float, int32_t)
    ORC_CASE(0, 1, float, int64_t)
    ORC_CASE(1, 0, double, int32_t)
    ORC_CASE(1, 1, double, int64_t)
    ORC_CASE(2, 0, c128, int32_t)
    ORC_CASE(2, 1, c128, int64_t)
#undef ORC_CASE
    if (J.plans) orc_plans_free(J.plans);
    i64 nnz = 0;
    double yn = 0.0;
    for (i64 r = 0; r < P; ++r) nnz += J.nnz[(size_t)r], yn += J.ynorm2[(size_t)r];
    if (nnz_out) *nnz_out = nnz;
    if (ynorm2_out) *ynorm2_out = yn;
    return rc;
}
