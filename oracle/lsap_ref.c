/* CPU restatement of the square linear-sum-assignment algorithm of csrc/traj_stats.cu::lsap_kernel -- TEST INFRASTRUCTURE
 * ONLY (oracle/: imported by tests/ only, never by the product path).
 *
 * The reference calls scipy.optimize.linear_sum_assignment (calc_statistics.py:60), a third-party dependency whose
 * source is not under /root/reference (version unpinned there).  Its published algorithm is the shortest augmenting
 * path method of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE T-AES 52(4), 2016: for
 * every row, grow a shortest-path tree over the columns (dual variables u, v), scanning the columns still outside the
 * tree; the next column is the one with the lowest path cost, ties resolved in favour of an UNASSIGNED column
 * (the scan updates on `<`, or on `==` when the column is unassigned); then update the duals and flip the path.
 * This file restates that scan sequentially, in the order the kernel's warp reduction is built to reproduce
 * ("lowest value; among equal values the LAST unassigned column in scan order if any, else the FIRST column"), so that
 * tests/test_statistics.py can pin the restatement to scipy itself on the CPU -- assignments, not only costs, also on
 * tie-heavy matrices -- before the GPU test compares the kernel with scipy.
 *
 * Build: gcc -O2 -shared -fPIC oracle/lsap_ref.c -o oracle/liblsap_ref.so  (oracle/build_lsap.py, __graft_entry__.build()).
 */
#include <math.h>
#include <stdlib.h>

/* cost [n][n] row-major; col4row [n] out.  Returns 0, or 1 if infeasible. */
int lsap_ref_solve(const double* cost, int n, int* col4row) {
    double* u = (double*)calloc((size_t)n, sizeof(double));
    double* v = (double*)calloc((size_t)n, sizeof(double));
    double* sp = (double*)malloc((size_t)n * sizeof(double));
    int* path = (int*)malloc((size_t)n * sizeof(int));
    int* row4col = (int*)malloc((size_t)n * sizeof(int));
    int* remaining = (int*)malloc((size_t)n * sizeof(int));
    unsigned char* SR = (unsigned char*)malloc((size_t)n);
    unsigned char* SC = (unsigned char*)malloc((size_t)n);
    int infeasible = 0;
    for (int j = 0; j < n; ++j) { col4row[j] = -1; row4col[j] = -1; path[j] = -1; }
    for (int cur = 0; cur < n && !infeasible; ++cur) {
        for (int j = 0; j < n; ++j) { sp[j] = INFINITY; SR[j] = 0; SC[j] = 0; remaining[j] = n - 1 - j; }
        int num_remaining = n, i = cur, sink = -1;
        double min_val = 0.0;
        while (sink < 0) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                const double r = min_val + cost[(size_t)i * n + j] - u[i] - v[j];
                if (r < sp[j]) { path[j] = i; sp[j] = r; }
                if (sp[j] < lowest || (sp[j] == lowest && row4col[j] == -1)) { lowest = sp[j]; index = it; }
            }
            min_val = lowest;
            if (!(min_val < INFINITY)) { infeasible = 1; break; }
            const int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (infeasible) break;
        u[cur] += min_val;
        for (int r = 0; r < n; ++r)
            if (SR[r] && r != cur) u[r] += min_val - sp[col4row[r]];
        for (int j = 0; j < n; ++j)
            if (SC[j]) v[j] -= min_val - sp[j];
        int j = sink;
        while (1) {
            const int r = path[j];
            row4col[j] = r;
            const int prev = col4row[r];
            col4row[r] = j;
            j = prev;
            if (r == cur) break;
        }
    }
    free(u); free(v); free(sp); free(path); free(row4col); free(remaining); free(SR); free(SC);
    return infeasible;
}
