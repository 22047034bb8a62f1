// Sample-set statistics of calc_statistics.py (reference calc_statistics.py:7-66): the 1-nearest-neighbour
// two-sample test and the earth mover's distance between K real and K generated trajectories of every pedestrian.
//
// The reference fills an n x n distance matrix per pedestrian with a Python double loop (O(n^2 T) interpreter steps),
// then np.argmin per row (1-NN) or scipy.optimize.linear_sum_assignment (EMD).  Here every pedestrian of every dump
// file is one independent problem of ONE launch:
//   traj_nn1_kernel   one warp per (pedestrian, sample): distances to all other samples of the mixed set, first-index
//                     argmin (np.argmin), classification counted into 4 integers with atomics (integer => deterministic)
//   traj_emd_cost_kernel  the n x n cost matrix in fp64, INCLUDING the reference's quirk: `D[ii, jj], D[jj, ii] = dij, dij`
//                     (calc_statistics.py:58) mirrors every entry, and the later write wins, so the matrix the solver
//                     sees is  D[a][b] = C[max(a,b)][min(a,b)],  C[a][b] = d(real a, fake b)
//   lsap_kernel       one warp per problem: shortest-augmenting-path assignment (Crouse 2016, the algorithm behind
//                     scipy's linear_sum_assignment), lane-parallel over the columns still outside the tree, with the
//                     sequential scan's tie-breaking rule reproduced in the warp reduction so that the ASSIGNMENT, not
//                     only its cost, equals scipy's.
//
// Distances follow numpy's arithmetic of `np.mean(np.sqrt(np.sum(np.power(diff, 2), 1)))` in the input dtype step by step
// (x*x, x2 + y2, correctly rounded sqrt, numpy's pairwise summation order, one division), with contraction into FMAs
// disabled through the _rn intrinsics: the matrices are bit-identical to the reference's, hence identical argmins.
#include "sw_common.cuh"

namespace sw {

template <typename T> struct Rn;
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

// || a_t - b_t ||_2 of one time step, numpy order: (dx*dx) + (dy*dy), then sqrt
template <typename T>
__device__ __forceinline__ T step_dist(const T* __restrict__ a, const T* __restrict__ b, int t) {
    const T dx = Rn<T>::sub(a[2 * t], b[2 * t]), dy = Rn<T>::sub(a[2 * t + 1], b[2 * t + 1]);
    return Rn<T>::sqrt(Rn<T>::add(Rn<T>::mul(dx, dx), Rn<T>::mul(dy, dy)));
}

// mean over t in [t0, t_len) of step_dist, summed the way numpy's pairwise_sum does for n <= 128 (one block):
// n < 8: sequential from 0; otherwise 8 running accumulators over the multiple-of-8 prefix, combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the tail sequentially.
template <typename T>
__device__ T traj_dist(const T* __restrict__ a, const T* __restrict__ b, int t0, int t_len) {
    const int n = t_len - t0;
    T res;
    if (n < 8) {
        res = (T)0;
        for (int i = 0; i < n; ++i) res = Rn<T>::add(res, step_dist(a, b, t0 + i));
    } else {
        T r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = step_dist(a, b, t0 + k);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = Rn<T>::add(r[k], step_dist(a, b, t0 + i + k));
        res = Rn<T>::add(Rn<T>::add(Rn<T>::add(r[0], r[1]), Rn<T>::add(r[2], r[3])),
                         Rn<T>::add(Rn<T>::add(r[4], r[5]), Rn<T>::add(r[6], r[7])));
        for (; i < n; ++i) res = Rn<T>::add(res, step_dist(a, b, t0 + i));
    }
    return Rn<T>::div(res, (T)n);
}

// samples [n][n_ped][t_len][2]
template <typename T>
__device__ __forceinline__ const T* sample_ptr(const T* reals, const T* fakes, int n_reals, int m, int ped, int n_ped, int t_len) {
    const T* base = m < n_reals ? reals : fakes;
    const int idx = m < n_reals ? m : m - n_reals;
    return base + ((size_t)idx * n_ped + ped) * t_len * 2;
}

// calc_statistics.py:7-46.  grid (ceil(n_mixed / 4), n_ped), 128 threads: warp = one row ii of pedestrian kk.
// counts[4] += (Real_pos, Real_neg, Fake_pos, Fake_neg)
template <typename T>
__global__ void traj_nn1_kernel(const T* __restrict__ reals, const T* __restrict__ fakes, int n_reals, int n_fakes, int n_ped,
                                int t_len, int obsv_len, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31, ii = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), ped = blockIdx.y;
    const int n_mixed = n_reals + n_fakes;
    if (ii >= n_mixed) return;
    const T* a = sample_ptr(reals, fakes, n_reals, ii, ped, n_ped, t_len);
    T best = (T)0;
    int best_j = -1;
    for (int jj = lane; jj < n_mixed; jj += 32) {
        // D is initialised to 1000 and only i != j entries are overwritten (:20, :29-33)
        const T d = jj == ii ? (T)1000 : traj_dist(a, sample_ptr(reals, fakes, n_reals, jj, ped, n_ped, t_len), obsv_len, t_len);
        if (best_j < 0 || d < best) { best = d; best_j = jj; }        // ascending jj per lane: first minimum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (oj >= 0 && (best_j < 0 || ob < best || (ob == best && oj < best_j))) { best = ob; best_j = oj; }
    }
    if (lane == 0) {
        const bool real_i = ii < n_reals, real_nn = best_j < n_reals;
        atomicAdd(counts + (real_i ? (real_nn ? 0 : 1) : (real_nn ? 3 : 2)), 1);
    }
}

// calc_statistics.py:54-58 with the mirrored write: D[ped][a][b] = C[max(a,b)][min(a,b)] (fp64 like np.ones(...)*1000)
template <typename T>
__global__ void traj_emd_cost_kernel(const T* __restrict__ reals, const T* __restrict__ fakes, int n, int n_ped, int t_len,
                                     int obsv_len, double* __restrict__ cost) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y, ped = blockIdx.z;
    if (b >= n) return;
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    const T d = traj_dist(reals + ((size_t)hi * n_ped + ped) * t_len * 2, fakes + ((size_t)lo * n_ped + ped) * t_len * 2,
                          obsv_len, t_len);
    cost[((size_t)ped * n + a) * n + b] = (double)d;
}

// Square linear sum assignment, one warp per problem (blockIdx.x).  cost [P][n][n] fp64 row-major; col4row [P][n] out.
// Shared memory: u, v, shortest (fp64 [n] each) | path, col4row, row4col, remaining (int [n] each) | SR, SC (bytes [n]).
__global__ void __launch_bounds__(32) lsap_kernel(const double* __restrict__ cost, int n, int* __restrict__ col4row_out,
                                                  int* __restrict__ status) {
    extern __shared__ double lsap_smem[];
    double* u = lsap_smem;
    double* v = u + n;
    double* sp = v + n;
    int* path = reinterpret_cast<int*>(sp + n);
    int* col4row = path + n;
    int* row4col = col4row + n;
    int* remaining = row4col + n;
    unsigned char* SR = reinterpret_cast<unsigned char*>(remaining + n);
    unsigned char* SC = SR + n;
    const int lane = threadIdx.x;
    const double* C = cost + (size_t)blockIdx.x * n * n;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int j = lane; j < n; j += 32) { u[j] = 0.0; v[j] = 0.0; col4row[j] = -1; row4col[j] = -1; path[j] = -1; }
    __syncwarp();
    bool infeasible = false;
    for (int cur = 0; cur < n && !infeasible; ++cur) {
        for (int j = lane; j < n; j += 32) { sp[j] = INF; SR[j] = 0; SC[j] = 0; remaining[j] = n - 1 - j; }
        __syncwarp();
        int num_remaining = n, i = cur, sink = -1;
        double min_val = 0.0;
        while (sink < 0) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            double bv = INF;
            int bit = -1;
            bool bun = false;
            for (int it = lane; it < num_remaining; it += 32) {
                const int j = remaining[it];
                const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, C[(size_t)i * n + j]), ui), v[j]);
                double s = sp[j];
                if (r < s) { path[j] = i; sp[j] = r; s = r; }
                const bool un = row4col[j] < 0;
                if (bit < 0 || s < bv || (s == bv && un)) { bv = s; bit = it; bun = un; }
            }
            // warp combine = the sequential scan's rule: lowest value; among equal values the LAST unassigned column in
            // scan order if any, else the FIRST column
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oit = __shfl_xor_sync(0xffffffffu, bit, o);
                const bool oun = __shfl_xor_sync(0xffffffffu, (int)bun, o) != 0;
                bool take = false;
                if (oit >= 0) {
                    if (bit < 0 || ov < bv) take = true;
                    else if (ov == bv) {
                        if (oun != bun) take = oun;
                        else take = oun ? (oit > bit) : (oit < bit);
                    }
                }
                if (take) { bv = ov; bit = oit; bun = oun; }
            }
            min_val = bv;
            if (!(min_val < INF)) { infeasible = true; break; }
            const int j = remaining[bit];
            const int r4c = row4col[j];
            if (r4c < 0) sink = j; else i = r4c;
            __syncwarp();
            if (lane == 0) { SC[j] = 1; remaining[bit] = remaining[num_remaining - 1]; }
            --num_remaining;
            __syncwarp();
        }
        if (infeasible) break;
        // dual variables
        if (lane == 0) u[cur] = __dadd_rn(u[cur], min_val);
        for (int r = lane; r < n; r += 32)
            if (SR[r] && r != cur) u[r] = __dadd_rn(u[r], __dsub_rn(min_val, sp[col4row[r]]));
        for (int j = lane; j < n; j += 32)
            if (SC[j]) v[j] = __dsub_rn(v[j], __dsub_rn(min_val, sp[j]));
        __syncwarp();
        // augment along the path back to the current row
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int r = path[j];
                row4col[j] = r;
                const int prev = col4row[r];
                col4row[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    for (int r = lane; r < n; r += 32) col4row_out[(size_t)blockIdx.x * n + r] = infeasible ? -1 : col4row[r];
    if (lane == 0 && infeasible) atomicExch(status, 1);
}

template <typename T>
int launch_nn1(const void* reals, const void* fakes, int n_reals, int n_fakes, int n_ped, int t_len, int obsv_len, int* counts,
               cudaStream_t st) {
    const int n_mixed = n_reals + n_fakes;
    dim3 grid((n_mixed + 3) / 4, n_ped);
    traj_nn1_kernel<T><<<grid, 128, 0, st>>>((const T*)reals, (const T*)fakes, n_reals, n_fakes, n_ped, t_len, obsv_len, counts);
    return 0;
}

template <typename T>
int launch_emd_cost(const void* reals, const void* fakes, int n, int n_ped, int t_len, int obsv_len, double* cost, cudaStream_t st) {
    dim3 grid((n + 127) / 128, n, n_ped);
    traj_emd_cost_kernel<T><<<grid, 128, 0, st>>>((const T*)reals, (const T*)fakes, n, n_ped, t_len, obsv_len, cost);
    return 0;
}

}  // namespace sw

static bool stats_args_ok(const void* reals, const void* fakes, int n_reals, int n_fakes, int n_ped, int t_len, int obsv_len,
                          int dtype_bytes) {
    return reals && fakes && n_reals > 0 && n_fakes > 0 && n_ped > 0 && t_len > 0 && obsv_len >= 0 && obsv_len < t_len &&
           (dtype_bytes == 4 || dtype_bytes == 8);
}

extern "C" int sw_traj_nn1_counts(const void* reals, const void* fakes, int dtype_bytes, int n_reals, int n_fakes, int n_ped,
                                  int t_len, int obsv_len, int* counts, void* stream) {
    if (!stats_args_ok(reals, fakes, n_reals, n_fakes, n_ped, t_len, obsv_len, dtype_bytes) || !counts) return SW_ERR_ARG;
    if (t_len - obsv_len > 128 || n_ped > 65535) return SW_ERR_UNSUPPORTED;   // one numpy pairwise block; gridDim.y
    cudaStream_t st = (cudaStream_t)stream;
    SW_CUDA_TRY(cudaMemsetAsync(counts, 0, 4 * sizeof(int), st));
    if (dtype_bytes == 4) sw::launch_nn1<float>(reals, fakes, n_reals, n_fakes, n_ped, t_len, obsv_len, counts, st);
    else sw::launch_nn1<double>(reals, fakes, n_reals, n_fakes, n_ped, t_len, obsv_len, counts, st);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_traj_emd_cost(const void* reals, const void* fakes, int dtype_bytes, int n, int n_ped, int t_len,
                                int obsv_len, double* cost, void* stream) {
    if (!stats_args_ok(reals, fakes, n, n, n_ped, t_len, obsv_len, dtype_bytes) || !cost) return SW_ERR_ARG;
    if (t_len - obsv_len > 128 || n_ped > 65535 || n > 65535) return SW_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype_bytes == 4) sw::launch_emd_cost<float>(reals, fakes, n, n_ped, t_len, obsv_len, cost, st);
    else sw::launch_emd_cost<double>(reals, fakes, n, n_ped, t_len, obsv_len, cost, st);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}

extern "C" int sw_lsap_smem_bytes(int n) { return n <= 0 ? 0 : n * (3 * 8 + 4 * 4 + 2); }

extern "C" int sw_lsap_solve(const double* cost, int n, int n_problems, int* col4row, int* status, void* stream) {
    if (!cost || !col4row || !status || n <= 0 || n_problems <= 0) return SW_ERR_ARG;
    const int smem = sw_lsap_smem_bytes(n);
    if (smem > 200 * 1024) return SW_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    SW_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    SW_SET_MAX_SMEM(sw::lsap_kernel, smem);
    sw::lsap_kernel<<<n_problems, 32, smem, st>>>(cost, n, col4row, status);
    SW_CUDA_TRY(cudaGetLastError());
    return SW_OK;
}
