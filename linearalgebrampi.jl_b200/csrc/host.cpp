// host.cpp — host-side structure of the hot path: partitions, column compression, the VectorPlan and TransposePlan
// index logic.  Pure C++ (no device).  The reference does this work in Julia; the algorithms here produce the same
// arrays by different means (segment arithmetic on the sorted ghost map, bitmap rank for compression, counting sort
// for the transpose) — see the file:line citations on each function and DESIGN.md §3.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "common.h"

namespace hpcla {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}
int host_threads(i64 work_items) {
    if (work_items < (i64)1 << 20) return 1;
    unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? (int)hc : 4;
    if (t > 32) t = 32;
    return t;
}
void parallel_for_threads(int nthreads, void (*fn)(int, int, void*), void* arg) {
    if (nthreads <= 1) {
        fn(0, 1, arg);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(fn, t, nthreads, arg);
    for (auto& x : th) x.join();
}

template <class F>
static void pfor(i64 n, F f) {  // f(begin, end) on disjoint chunks
    int nt = host_threads(n);
    if (nt <= 1) {
        f((i64)0, n);
        return;
    }
    std::vector<std::thread> th;
    i64 chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        i64 b = t * chunk, e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=]() { f(b, e); });
    }
    for (auto& x : th) x.join();
}

// Sorted-unique + rank of a multiset of 1-based ids in [1, universe] with a bitmap and per-word prefix counts.
struct BitRank {
    i64 universe = 0;
    std::vector<uint64_t> bits;
    std::vector<i64> prefix;  // set bits before word w
    void init(i64 u) {
        universe = u;
        bits.assign((size_t)((u + 64) / 64), 0);
    }
    inline void mark_atomic(i64 g) {  // g 1-based
        uint64_t m = (uint64_t)1 << ((g - 1) & 63);
        uint64_t* w = &bits[(size_t)((g - 1) >> 6)];
        if (!(__atomic_load_n(w, __ATOMIC_RELAXED) & m)) __atomic_fetch_or(w, m, __ATOMIC_RELAXED);
    }
    i64 finalize() {
        prefix.resize(bits.size() + 1);
        i64 acc = 0;
        for (size_t w = 0; w < bits.size(); ++w) {
            prefix[w] = acc;
            acc += __builtin_popcountll(bits[w]);
        }
        prefix[bits.size()] = acc;
        return acc;
    }
    inline i64 rank1(i64 g) const {  // 1-based position of g among the marked ids
        size_t w = (size_t)((g - 1) >> 6);
        unsigned b = (unsigned)((g - 1) & 63);
        uint64_t below = b ? (bits[w] & (((uint64_t)1 << b) - 1)) : 0;
        return prefix[w] + __builtin_popcountll(below) + 1;
    }
    void enumerate(i64* out) const {
        pfor((i64)bits.size(), [&](i64 wb, i64 we) {
            for (i64 w = wb; w < we; ++w) {
                uint64_t x = bits[(size_t)w];
                i64 k = prefix[(size_t)w];
                while (x) {
                    int b = __builtin_ctzll(x);
                    out[k++] = w * 64 + b + 1;
                    x &= x - 1;
                }
            }
        });
    }
};

template <class Ti>
static int compress_impl(i64 nnz, const Ti* g, i64 ncols_global, Ti* colval, i64* col_indices, i64* ncc_out) {
    BitRank br;
    br.init(ncols_global);
    std::atomic<int> bad(0);
    pfor(nnz, [&](i64 b, i64 e) {
        for (i64 k = b; k < e; ++k) {
            i64 c = (i64)g[k];
            if (c < 1 || c > ncols_global) {
                bad.store(1);
                return;
            }
            br.mark_atomic(c);
        }
    });
    if (bad.load()) return fail(HPCLA_ERR_ARG, "hpcla_compress_columns: a column index lies outside [1, %lld]", (long long)ncols_global);
    i64 ncc = br.finalize();
    br.enumerate(col_indices);
    pfor(nnz, [&](i64 b, i64 e) {
        for (i64 k = b; k < e; ++k) colval[k] = (Ti)br.rank1((i64)g[k]);
    });
    *ncc_out = ncc;
    return HPCLA_OK;
}

}  // namespace hpcla

using namespace hpcla;

extern "C" int hpcla_abi_version(void) { return HPCLA_ABI_VERSION; }
extern "C" const char* hpcla_last_error(void) { return last_error_ref().c_str(); }

// src/HPCLinearAlgebra.jl:279-289
extern "C" int hpcla_uniform_partition(int64_t n, int nranks, int64_t* out) {
    if (n < 0 || nranks < 1 || !out) return fail(HPCLA_ERR_ARG, "hpcla_uniform_partition: bad arguments");
    i64 q = n / nranks, rem = n % nranks;
    out[0] = 1;
    for (int r = 0; r < nranks; ++r) out[r + 1] = out[r] + q + (r < rem ? 1 : 0);
    return HPCLA_OK;
}

// VectorRepartitionPlan(x, p) — src/vectors.jl:519-616.  Both partitions are known on every rank, so the plan is a pure
// function (the reference's Alltoall of counts, :556-557, only tells a rank what it could compute itself).  Ranks
// whose range in the other partition meets mine form one consecutive run: found by bisection, then walked.
static int first_rank_reaching(const int64_t* part, int nranks, i64 g) {  // first rank r with part[r+1] > g
    int lo = 0, hi = nranks;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (part[mid + 1] > g) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

extern "C" int hpcla_repartition_plan(int rank, int nranks, const int64_t* old_partition, const int64_t* new_partition, int64_t* n_send_out,
                                      int64_t* send_rank_ids, int64_t* send_first, int64_t* send_count, int64_t* n_recv_out, int64_t* recv_rank_ids,
                                      int64_t* recv_count, int64_t* recv_offset, int64_t* local_out /* [3]: src first, count, dst offset */,
                                      int64_t* result_local_size_out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !old_partition || !new_partition || !n_send_out || !n_recv_out || !local_out || !result_local_size_out)
        return fail(HPCLA_ERR_ARG, "hpcla_repartition_plan: bad arguments");
    if (old_partition[0] != 1 || new_partition[0] != 1 || old_partition[nranks] != new_partition[nranks])
        return fail(HPCLA_ERR_ARG, "hpcla_repartition_plan: partitions must start at 1 and cover the same %lld elements", (long long)(old_partition[nranks] - 1));
    for (int r = 0; r < nranks; ++r)
        if (old_partition[r + 1] < old_partition[r] || new_partition[r + 1] < new_partition[r]) return fail(HPCLA_ERR_ARG, "hpcla_repartition_plan: partitions must be non-decreasing");
    const i64 src_b = old_partition[rank], src_e = old_partition[rank + 1];  // my elements today: [src_b, src_e)
    const i64 dst_b = new_partition[rank], dst_e = new_partition[rank + 1];  // my elements afterwards
    i64 ns = 0, nr = 0;
    local_out[0] = 1, local_out[1] = 0, local_out[2] = 0;  // local_src_range = 1:0, local_dst_offset = 0 (:571-572)
    if (src_e > src_b) {  // who receives my elements: ranks whose NEW range meets [src_b, src_e)
        for (int r = first_rank_reaching(new_partition, nranks, src_b); r < nranks && new_partition[r] < src_e; ++r) {
            const i64 b = std::max(src_b, new_partition[r]), e = std::min(src_e, new_partition[r + 1]);
            if (e <= b) continue;
            if (r == rank) {
                local_out[0] = b - src_b + 1;
                local_out[1] = e - b;
                local_out[2] = b - dst_b + 1;
            } else {
                send_rank_ids[ns] = r, send_first[ns] = b - src_b + 1, send_count[ns] = e - b;
                ++ns;
            }
        }
    }
    if (dst_e > dst_b) {  // who sends me elements: ranks whose OLD range meets [dst_b, dst_e)
        for (int r = first_rank_reaching(old_partition, nranks, dst_b); r < nranks && old_partition[r] < dst_e; ++r) {
            const i64 b = std::max(dst_b, old_partition[r]), e = std::min(dst_e, old_partition[r + 1]);
            if (e <= b || r == rank) continue;
            recv_rank_ids[nr] = r, recv_count[nr] = e - b, recv_offset[nr] = b - dst_b + 1;
            ++nr;
        }
    }
    *n_send_out = ns;
    *n_recv_out = nr;
    *result_local_size_out = dst_e - dst_b;
    return HPCLA_OK;
}

// src/sparse.jl:501 (col_indices) and :137-144 (compress_AT)
extern "C" int hpcla_compress_columns(int itype, int64_t nnz, const void* global_cols, int64_t ncols_global, void* colval_out,
                                      int64_t* col_indices_out, int64_t* ncc_out) {
    if (nnz < 0 || ncols_global < 0 || !ncc_out) return fail(HPCLA_ERR_ARG, "hpcla_compress_columns: bad arguments");
    if (nnz == 0) {
        *ncc_out = 0;
        return HPCLA_OK;
    }
    if (itype == HPCLA_I32) return compress_impl<int32_t>(nnz, (const int32_t*)global_cols, ncols_global, (int32_t*)colval_out, col_indices_out, ncc_out);
    if (itype == HPCLA_I64) return compress_impl<int64_t>(nnz, (const int64_t*)global_cols, ncols_global, (int64_t*)colval_out, col_indices_out, ncc_out);
    return fail(HPCLA_ERR_ARG, "hpcla_compress_columns: unknown index type %d", itype);
}

// ---------------------------------------------------------------------------------------------------------------
// VectorPlan(A, x) — src/sparse.jl:1875-1984.
// col_indices is sorted and every rank of x.partition owns a contiguous global range, so the per-owner lists of
// (:1886-1895) are contiguous segments of col_indices: found with one binary search per partition boundary.
// searchsortedlast semantics with repeated boundaries (empty ranks) and the clamp of :1892-1894 fall out of using
// lower_bound on the boundary values and giving the last rank everything above its start.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int hpcla_plan_begin(int rank, int nranks, const int64_t* col_indices, int64_t ncc, const int64_t* x_partition,
                                hpcla_planb** out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || ncc < 0 || !x_partition || !out || (ncc > 0 && !col_indices))
        return fail(HPCLA_ERR_ARG, "hpcla_plan_begin: bad arguments");
    for (int r = 0; r < nranks; ++r)
        if (x_partition[r + 1] < x_partition[r]) return fail(HPCLA_ERR_ARG, "hpcla_plan_begin: x partition is not non-decreasing");
    for (i64 k = 1; k < ncc; ++k)
        if (col_indices[k] <= col_indices[k - 1]) return fail(HPCLA_ERR_ARG, "hpcla_plan_begin: col_indices must be strictly ascending");
    if (ncc > 0 && col_indices[0] < x_partition[0]) return fail(HPCLA_ERR_ARG, "hpcla_plan_begin: col_indices[1] precedes the partition");
    hpcla_planb* pb = new hpcla_planb();
    pb->rank = rank;
    pb->nranks = nranks;
    pb->ncc = ncc;
    pb->col_indices.assign(col_indices, col_indices + ncc);
    pb->x_partition.assign(x_partition, x_partition + nranks + 1);
    pb->seg_start.resize(nranks + 1);
    const i64* ci = pb->col_indices.data();
    for (int o = 0; o < nranks; ++o) pb->seg_start[o] = (i64)(std::lower_bound(ci, ci + ncc, x_partition[o]) - ci);
    pb->seg_start[nranks] = ncc;
    pb->seg_start[0] = 0;
    *out = pb;
    return HPCLA_OK;
}
extern "C" int hpcla_planb_counts(const hpcla_planb* pb, int64_t* counts) {
    if (!pb || !counts) return fail(HPCLA_ERR_ARG, "hpcla_planb_counts: null");
    for (int o = 0; o < pb->nranks; ++o) counts[o] = pb->seg_start[o + 1] - pb->seg_start[o];
    return HPCLA_OK;
}
extern "C" int hpcla_planb_requests(const hpcla_planb* pb, int owner, int64_t* out) {
    if (!pb || owner < 0 || owner >= pb->nranks) return fail(HPCLA_ERR_ARG, "hpcla_planb_requests: bad owner");
    i64 b = pb->seg_start[owner], e = pb->seg_start[owner + 1];
    if (e > b) std::memcpy(out, pb->col_indices.data() + b, (size_t)(e - b) * sizeof(i64));
    return HPCLA_OK;
}
extern "C" int hpcla_plan_finish(hpcla_planb* pb, const int64_t* recv_counts, const int64_t* const* recv_lists, hpcla_plan** out) {
    if (!pb || !recv_counts || !out) return fail(HPCLA_ERR_ARG, "hpcla_plan_finish: null");
    hpcla_plan* pl = new hpcla_plan();
    pl->rank = pb->rank;
    pl->nranks = pb->nranks;
    pl->n_gathered = pb->ncc;
    const int me = pb->rank;
    const i64 x0 = pb->x_partition[me];
    const i64 nloc = pb->x_partition[me + 1] - x0;
    pl->n_x_local = nloc;
    for (int o = 0; o < pb->nranks; ++o) {
        i64 b = pb->seg_start[o], e = pb->seg_start[o + 1];
        if (e == b) continue;
        if (o == me) {  // :1947-1953
            pl->local_src.resize((size_t)(e - b));
            pl->local_dst.resize((size_t)(e - b));
            for (i64 k = b; k < e; ++k) {
                pl->local_src[(size_t)(k - b)] = pb->col_indices[(size_t)k] - x0 + 1;
                pl->local_dst[(size_t)(k - b)] = k + 1;
            }
        } else {  // :1908-1918
            pl->recv_rank_ids.push_back(o);
            std::vector<i64> perm((size_t)(e - b));
            for (i64 k = b; k < e; ++k) perm[(size_t)(k - b)] = k + 1;
            pl->recv_perm.push_back(std::move(perm));
        }
    }
    for (int q = 0; q < pb->nranks; ++q) {  // :1925-1944
        if (q == me || recv_counts[q] <= 0) continue;
        if (!recv_lists || !recv_lists[q]) {
            delete pl;
            return fail(HPCLA_ERR_ARG, "hpcla_plan_finish: missing request list from rank %d", q);
        }
        pl->send_rank_ids.push_back(q);
        std::vector<i64> loc((size_t)recv_counts[q]);
        for (i64 k = 0; k < recv_counts[q]; ++k) {
            i64 li = recv_lists[q][k] - x0 + 1;
            if (li < 1 || li > nloc) {
                delete pl;
                return fail(HPCLA_ERR_ARG, "hpcla_plan_finish: rank %d requested global index %lld which rank %d does not own",
                            q, (long long)recv_lists[q][k], me);
            }
            loc[(size_t)k] = li;
        }
        pl->send_indices.push_back(std::move(loc));
    }
    delete pb;
    *out = pl;
    return HPCLA_OK;
}

template <class Ti>
static void widen(const void* src, i64 n, std::vector<i64>& dst) {
    dst.resize((size_t)n);
    const Ti* s = (const Ti*)src;
    for (i64 k = 0; k < n; ++k) dst[(size_t)k] = (i64)s[k];
}
static void widen_any(int itype, const void* src, i64 n, std::vector<i64>& dst) {
    if (itype == HPCLA_I32) widen<int32_t>(src, n, dst);
    else widen<int64_t>(src, n, dst);
}

extern "C" int hpcla_plan_import(int rank, int nranks, int itype, int64_t n_gathered, int64_t n_x_local, int64_t n_send,
                                 const int64_t* send_rank_ids, const int64_t* send_lens, const void* const* send_indices,
                                 int64_t n_recv, const int64_t* recv_rank_ids, const int64_t* recv_lens,
                                 const void* const* recv_perm, int64_t n_local, const void* local_src, const void* local_dst,
                                 hpcla_plan** out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !out || n_send < 0 || n_recv < 0 || n_local < 0 || n_gathered < 0)
        return fail(HPCLA_ERR_ARG, "hpcla_plan_import: bad arguments");
    if (itype != HPCLA_I32 && itype != HPCLA_I64) return fail(HPCLA_ERR_ARG, "hpcla_plan_import: unknown index type");
    hpcla_plan* pl = new hpcla_plan();
    pl->rank = rank;
    pl->nranks = nranks;
    pl->n_gathered = n_gathered;
    pl->n_x_local = n_x_local;
    for (i64 i = 0; i < n_send; ++i) {
        pl->send_rank_ids.push_back(send_rank_ids[i]);
        pl->send_indices.emplace_back();
        widen_any(itype, send_indices[i], send_lens[i], pl->send_indices.back());
    }
    for (i64 i = 0; i < n_recv; ++i) {
        pl->recv_rank_ids.push_back(recv_rank_ids[i]);
        pl->recv_perm.emplace_back();
        widen_any(itype, recv_perm[i], recv_lens[i], pl->recv_perm.back());
    }
    widen_any(itype, local_src, n_local, pl->local_src);
    widen_any(itype, local_dst, n_local, pl->local_dst);
    // validate ranges: a wrong plan must fail here, not fault on the device
    auto in_range = [](const std::vector<i64>& v, i64 hi) {
        for (i64 x : v)
            if (x < 1 || x > hi) return false;
        return true;
    };
    bool ok = in_range(pl->local_dst, n_gathered) && (n_x_local < 0 || in_range(pl->local_src, n_x_local));
    for (auto& v : pl->recv_perm) ok = ok && in_range(v, n_gathered);
    for (auto& v : pl->send_indices) ok = ok && (n_x_local < 0 || in_range(v, n_x_local));
    for (i64 r : pl->send_rank_ids) ok = ok && r >= 0 && r < nranks && r != rank;
    for (i64 r : pl->recv_rank_ids) ok = ok && r >= 0 && r < nranks && r != rank;
    if (!ok) {
        delete pl;
        return fail(HPCLA_ERR_ARG, "hpcla_plan_import: an index or rank id is out of range");
    }
    *out = pl;
    return HPCLA_OK;
}

static const std::vector<i64>* plan_field(const hpcla_plan* p, int field, i64 slot) {
    switch (field) {
        case 0: return &p->send_rank_ids;
        case 1: return &p->recv_rank_ids;
        case 2: return &p->local_src;
        case 3: return &p->local_dst;
        case 4: return (slot >= 0 && slot < (i64)p->send_indices.size()) ? &p->send_indices[(size_t)slot] : nullptr;
        case 5: return (slot >= 0 && slot < (i64)p->recv_perm.size()) ? &p->recv_perm[(size_t)slot] : nullptr;
    }
    return nullptr;
}
extern "C" int hpcla_plan_len(const hpcla_plan* plan, int field, int64_t slot, int64_t* len_out) {
    const std::vector<i64>* v = plan ? plan_field(plan, field, slot) : nullptr;
    if (!v || !len_out) return fail(HPCLA_ERR_ARG, "hpcla_plan_len: bad field/slot");
    *len_out = (i64)v->size();
    return HPCLA_OK;
}
extern "C" int hpcla_plan_get(const hpcla_plan* plan, int field, int64_t slot, int64_t* out) {
    const std::vector<i64>* v = plan ? plan_field(plan, field, slot) : nullptr;
    if (!v) return fail(HPCLA_ERR_ARG, "hpcla_plan_get: bad field/slot");
    if (!v->empty()) std::memcpy(out, v->data(), v->size() * sizeof(i64));
    return HPCLA_OK;
}
extern "C" int hpcla_plan_n_gathered(const hpcla_plan* plan, int64_t* n_out) {
    if (!plan || !n_out) return fail(HPCLA_ERR_ARG, "hpcla_plan_n_gathered: null");
    *n_out = plan->n_gathered;
    return HPCLA_OK;
}
extern "C" void hpcla_plan_destroy(hpcla_plan* plan) { delete plan; }

// ---------------------------------------------------------------------------------------------------------------
// TransposePlan(A) / execute_plan!(plan, A) — src/sparse.jl:1551-1744, 1756-1829.
// The reference sorts (j, i, src_rank, src_idx) tuples (:1655).  Here: nonzeros are scanned in stored order
// (i ascending, then j ascending), bucketed by destination; the receiver walks the sources in ascending rank order
// (ascending, disjoint i ranges) and does a stable counting sort on j — same result, O(nnz).
// ---------------------------------------------------------------------------------------------------------------
struct hpcla_tb {
    int rank = 0, nranks = 1, dtype = 0, itype = 0;
    size_t es = 0;
    std::vector<i64> row_partition, col_partition;
    std::vector<i64> counts;                  // per destination
    std::vector<std::vector<i64>> pairs;      // per destination: (j, i) interleaved
    std::vector<std::vector<char>> vals;      // per destination
    // result
    bool finished = false;
    i64 nrows = 0, nnz = 0, ncc = 0;
    std::vector<i64> rowptr, colval_compressed, col_indices;
    std::vector<char> nzval;
};

template <class Ti>
static int tb_begin_impl(hpcla_tb* tb, const Ti* rowptr, const Ti* colval, const i64* col_indices, const char* nzval) {
    const int P = tb->nranks, me = tb->rank;
    const i64 row0 = tb->row_partition[me];
    const i64 nrows = tb->row_partition[me + 1] - row0;
    const i64* cp = tb->col_partition.data();
    const size_t es = tb->es;
    tb->counts.assign(P, 0);
    auto dest_of = [&](i64 j) {
        int d = (int)(std::upper_bound(cp, cp + P + 1, j) - cp) - 1;  // searchsortedlast(col_partition, j) - 1 (:1575)
        return d >= P ? P - 1 : d;
    };
    const i64 nnz = nrows > 0 ? (i64)rowptr[nrows] - 1 : 0;
    std::vector<int> dest((size_t)nnz);
    for (i64 k = 0; k < nnz; ++k) {
        i64 j = col_indices[(i64)colval[k] - 1];
        int d = dest_of(j);
        if (d < 0) return fail(HPCLA_ERR_ARG, "hpcla_transpose_begin: column %lld precedes col_partition", (long long)j);
        dest[(size_t)k] = d;
        tb->counts[d] += 1;
    }
    tb->pairs.resize(P);
    tb->vals.resize(P);
    std::vector<i64> cur(P, 0);
    for (int d = 0; d < P; ++d) {
        tb->pairs[d].resize((size_t)(2 * tb->counts[d]));
        tb->vals[d].resize((size_t)tb->counts[d] * es);
    }
    for (i64 li = 0; li < nrows; ++li) {
        i64 gi = row0 + li;
        for (i64 k = (i64)rowptr[li] - 1; k < (i64)rowptr[li + 1] - 1; ++k) {
            int d = dest[(size_t)k];
            i64 c = cur[d]++;
            tb->pairs[d][(size_t)(2 * c)] = col_indices[(i64)colval[k] - 1];  // row of A^T (:1598)
            tb->pairs[d][(size_t)(2 * c + 1)] = gi;                           // col of A^T (:1599)
            std::memcpy(tb->vals[d].data() + (size_t)c * es, nzval + (size_t)k * es, es);
        }
    }
    return HPCLA_OK;
}

extern "C" int hpcla_transpose_begin(int rank, int nranks, int dtype, int itype, const int64_t* row_partition,
                                     const int64_t* col_partition, const void* h_rowptr, const void* h_colval,
                                     const int64_t* h_col_indices, const void* h_nzval, hpcla_tb** out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !row_partition || !col_partition || !h_rowptr || !out)
        return fail(HPCLA_ERR_ARG, "hpcla_transpose_begin: bad arguments");
    if (!dtype_size(dtype) || !itype_size(itype)) return fail(HPCLA_ERR_ARG, "hpcla_transpose_begin: unknown dtype/itype");
    hpcla_tb* tb = new hpcla_tb();
    tb->rank = rank;
    tb->nranks = nranks;
    tb->dtype = dtype;
    tb->itype = itype;
    tb->es = dtype_size(dtype);
    tb->row_partition.assign(row_partition, row_partition + nranks + 1);
    tb->col_partition.assign(col_partition, col_partition + nranks + 1);
    int rc = itype == HPCLA_I32 ? tb_begin_impl<int32_t>(tb, (const int32_t*)h_rowptr, (const int32_t*)h_colval, h_col_indices, (const char*)h_nzval)
                                : tb_begin_impl<int64_t>(tb, (const int64_t*)h_rowptr, (const int64_t*)h_colval, h_col_indices, (const char*)h_nzval);
    if (rc != HPCLA_OK) {
        delete tb;
        return rc;
    }
    *out = tb;
    return HPCLA_OK;
}
extern "C" int hpcla_tb_counts(const hpcla_tb* tb, int64_t* counts_out) {
    if (!tb || !counts_out) return fail(HPCLA_ERR_ARG, "hpcla_tb_counts: null");
    std::copy(tb->counts.begin(), tb->counts.end(), counts_out);
    return HPCLA_OK;
}
extern "C" int hpcla_tb_message(const hpcla_tb* tb, int dest, int64_t* pairs_out, void* vals_out) {
    if (!tb || dest < 0 || dest >= tb->nranks) return fail(HPCLA_ERR_ARG, "hpcla_tb_message: bad destination");
    if (tb->pairs.empty()) return fail(HPCLA_ERR_STATE, "hpcla_tb_message: messages were released by finish");
    if (tb->counts[dest] > 0) {
        std::memcpy(pairs_out, tb->pairs[dest].data(), tb->pairs[dest].size() * sizeof(i64));
        std::memcpy(vals_out, tb->vals[dest].data(), tb->vals[dest].size());
    }
    return HPCLA_OK;
}
extern "C" int hpcla_transpose_finish(hpcla_tb* tb, const int64_t* recv_counts, const int64_t* const* recv_pairs,
                                      const void* const* recv_vals, int64_t* nrows_out, int64_t* nnz_out, int64_t* ncc_out) {
    if (!tb || !recv_counts) return fail(HPCLA_ERR_ARG, "hpcla_transpose_finish: null");
    const int P = tb->nranks, me = tb->rank;
    const size_t es = tb->es;
    const i64 at_row0 = tb->col_partition[me];
    const i64 nrows = tb->col_partition[me + 1] - at_row0;  // :1627-1629
    const i64 ncols_global = tb->row_partition[P] - 1;      // nrows_A (:1558)
    auto src_pairs = [&](int r) { return r == me ? tb->pairs[me].data() : recv_pairs[r]; };
    auto src_vals = [&](int r) { return r == me ? (const char*)tb->vals[me].data() : (const char*)recv_vals[r]; };
    auto src_count = [&](int r) { return r == me ? tb->counts[me] : recv_counts[r]; };
    i64 nnz = 0;
    for (int r = 0; r < P; ++r) {
        i64 c = src_count(r);
        if (c < 0 || (c > 0 && r != me && (!recv_pairs || !recv_pairs[r] || !recv_vals || !recv_vals[r])))
            return fail(HPCLA_ERR_ARG, "hpcla_transpose_finish: missing message from rank %d", r);
        nnz += c;
    }
    tb->rowptr.assign((size_t)nrows + 1, 0);
    for (int r = 0; r < P; ++r) {
        const i64* pr = src_pairs(r);
        for (i64 k = 0; k < src_count(r); ++k) {
            i64 lj = pr[2 * k] - at_row0;
            if (lj < 0 || lj >= nrows) return fail(HPCLA_ERR_ARG, "hpcla_transpose_finish: rank %d sent row %lld that rank %d does not own", r, (long long)pr[2 * k], me);
            tb->rowptr[(size_t)lj + 1] += 1;
        }
    }
    tb->rowptr[0] = 1;
    for (i64 j = 0; j < nrows; ++j) tb->rowptr[(size_t)j + 1] += tb->rowptr[(size_t)j];
    std::vector<i64> cursor(tb->rowptr.begin(), tb->rowptr.end() - 1);
    std::vector<i64> gcols((size_t)nnz);
    tb->nzval.resize((size_t)nnz * es);
    for (int r = 0; r < P; ++r) {  // ascending source rank == ascending i within every row j
        const i64* pr = src_pairs(r);
        const char* pv = src_vals(r);
        for (i64 k = 0; k < src_count(r); ++k) {
            i64 lj = pr[2 * k] - at_row0;
            i64 pos = cursor[(size_t)lj]++ - 1;
            gcols[(size_t)pos] = pr[2 * k + 1];
            std::memcpy(tb->nzval.data() + (size_t)pos * es, pv + (size_t)k * es, es);
        }
    }
    // result col_indices (:1723) + compression (:1802)
    tb->col_indices.resize((size_t)std::min<i64>(nnz, ncols_global));
    tb->colval_compressed.resize((size_t)nnz);
    i64 ncc = 0;
    if (nnz > 0) {
        int rc = compress_impl<int64_t>(nnz, gcols.data(), ncols_global, tb->colval_compressed.data(), tb->col_indices.data(), &ncc);
        if (rc != HPCLA_OK) return rc;
    }
    tb->col_indices.resize((size_t)ncc);
    tb->nrows = nrows;
    tb->nnz = nnz;
    tb->ncc = ncc;
    tb->finished = true;
    std::vector<std::vector<i64>>().swap(tb->pairs);
    std::vector<std::vector<char>>().swap(tb->vals);
    if (nrows_out) *nrows_out = nrows;
    if (nnz_out) *nnz_out = nnz;
    if (ncc_out) *ncc_out = ncc;
    return HPCLA_OK;
}
extern "C" int hpcla_tb_result(const hpcla_tb* tb, void* rowptr_out, void* colval_out, int64_t* col_indices_out, void* nzval_out) {
    if (!tb || !tb->finished) return fail(HPCLA_ERR_STATE, "hpcla_tb_result: finish has not run");
    if (tb->itype == HPCLA_I32) {
        int32_t* rp = (int32_t*)rowptr_out;
        int32_t* cv = (int32_t*)colval_out;
        if (tb->nnz + 1 > (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_tb_result: nnz does not fit Int32");
        for (size_t k = 0; k < tb->rowptr.size(); ++k) rp[k] = (int32_t)tb->rowptr[k];
        for (size_t k = 0; k < tb->colval_compressed.size(); ++k) cv[k] = (int32_t)tb->colval_compressed[k];
    } else {
        std::memcpy(rowptr_out, tb->rowptr.data(), tb->rowptr.size() * sizeof(i64));
        if (tb->nnz) std::memcpy(colval_out, tb->colval_compressed.data(), (size_t)tb->nnz * sizeof(i64));
    }
    if (tb->ncc) std::memcpy(col_indices_out, tb->col_indices.data(), (size_t)tb->ncc * sizeof(i64));
    if (tb->nnz) std::memcpy(nzval_out, tb->nzval.data(), tb->nzval.size());
    return HPCLA_OK;
}
extern "C" void hpcla_tb_destroy(hpcla_tb* tb) { delete tb; }
