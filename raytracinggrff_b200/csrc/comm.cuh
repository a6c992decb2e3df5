// Multi-GPU side of the map renderer: one process per GPU, the cube
// replicated, the image rows dealt to the ranks, ONE exchange at the end — the per-rank image slabs go
// to the root over NCCL (NVLink 5 / NVSwitch) and a kernel puts the rows at their final positions.
// Rays never interact, so nothing is exchanged while they are integrated.
//
// Replaces the reference's only parallel mode: contiguous ray chunks pickled to ProcessPoolExecutor
// workers and concatenated on the ray axis (script/resample_with_ray_tracing.py:42-61, :333-352).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded, if any, else the
// system's), so the library loads — and every single-GPU entry point works — without NCCL present.
#pragma once

#include <dlfcn.h>

#include "common.cuh"

namespace rtgrff {

// ---- row sharding ----------------------------------------------------------------------------------
// Rows are dealt round-robin in groups of kRowGroup adjacent rows (the height of the 4 x 8-pixel tile a
// warp walks, so a warp's 32 rays stay neighbours in the image) when every rank still gets at least 8
// groups, else row by row: disk-centre rays live much longer than limb rays, so contiguous bands would
// be badly balanced.  Group g belongs to rank g mod W.
constexpr int kRowGroup = 8;

__host__ __device__ inline int row_group_of(int n_rows, int world) { return n_rows >= 8 * kRowGroup * world ? kRowGroup : 1; }

// owner and index within the owner's slab of image row `row`
__host__ __device__ inline void row_owner(int row, int group, int world, int &rank, int &local)
{
    const int g = row / group;
    rank = g % world;
    local = (g / world) * group + row % group;
}

inline int rows_of_rank_count(int n_rows, int world, int rank)
{
    const int G = row_group_of(n_rows, world);
    int n = 0;
    for (int row = 0; row < n_rows; ++row) n += ((row / G) % world == rank);
    return n;
}

inline int max_rows_per_rank(int n_rows, int world)
{
    int m = 0;
    for (int r = 0; r < world; ++r) m = std::max(m, rows_of_rank_count(n_rows, world, r));
    return m;
}

// gathered [rank][plane][max_rows][n_cols] -> image [plane][n_rows][n_cols]
__global__ void place_rows_kernel(const double *__restrict__ gathered, double *__restrict__ image, int n_planes, int n_rows,
                                  int n_cols, int max_rows, int world, int group)
{
    const int64_t total = (int64_t)n_planes * n_rows * n_cols;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(q % n_cols);
        const int row = (int)((q / n_cols) % n_rows);
        const int plane = (int)(q / ((int64_t)n_cols * n_rows));
        int rank, local;
        row_owner(row, group, world, rank, local);
        image[q] = gathered[(((int64_t)rank * n_planes + plane) * max_rows + local) * n_cols + col];
    }
}

// ---- NCCL, bound at run time -----------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };     // ncclUniqueId
typedef void *NcclComm;                           // ncclComm_t
constexpr int kNcclInt8 = 0;                      // ncclInt8 / ncclChar

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

inline NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char *env = getenv("RTGRFF_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !n[0]) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return nullptr;
#define RT_SYM(field, name)                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));       \
    if (!api.field) { dlclose(h); return nullptr; }
    RT_SYM(GetUniqueId, "ncclGetUniqueId") RT_SYM(CommInitRank, "ncclCommInitRank") RT_SYM(CommDestroy, "ncclCommDestroy")
    RT_SYM(Send, "ncclSend") RT_SYM(Recv, "ncclRecv") RT_SYM(GroupStart, "ncclGroupStart") RT_SYM(GroupEnd, "ncclGroupEnd")
    RT_SYM(GetErrorString, "ncclGetErrorString") RT_SYM(GetVersion, "ncclGetVersion")
#undef RT_SYM
    api.handle = h;
    return &api;
}

}  // namespace rtgrff
