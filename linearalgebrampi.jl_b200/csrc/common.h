// common.h — internal declarations shared by the translation units of libhpcla_b200.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hpcla_b200.h"

typedef int64_t i64;

namespace hpcla {

std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

inline size_t dtype_size(int dtype) { return dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : dtype == HPCLA_C128 ? 16 : 0; }
inline size_t itype_size(int itype) { return itype == HPCLA_I32 ? 4 : itype == HPCLA_I64 ? 8 : 0; }

// number of worker threads for host-side passes over nnz-sized arrays
int host_threads(i64 work_items);
// run fn(t, nthreads) on nthreads std::threads (inline when nthreads == 1)
void parallel_for_threads(int nthreads, void (*fn)(int, int, void*), void* arg);

}  // namespace hpcla

// Host-side VectorPlan index fields (src/vectors.jl:229-251), 1-based Int64 values.
struct hpcla_plan {
    int rank = 0, nranks = 1;
    i64 n_gathered = 0;
    i64 n_x_local = -1;  // -1: unknown until bound
    std::vector<i64> send_rank_ids, recv_rank_ids;
    std::vector<std::vector<i64>> send_indices, recv_perm;
    std::vector<i64> local_src, local_dst;
};

struct hpcla_planb {
    int rank = 0, nranks = 1;
    i64 ncc = 0;
    const i64* col_indices_borrowed = nullptr;  // valid until finish
    std::vector<i64> col_indices;               // private copy (requests are read after begin returns)
    std::vector<i64> x_partition;
    std::vector<i64> seg_start;  // [nranks+1] 0-based offsets into col_indices per owner
};
