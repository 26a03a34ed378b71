// pd_patch.cu - builds the bicubic patches of pd_patch.h on the device (once per table and GPU).
#include <cmath>
#include <cstring>
#include <vector>

#include "pd_patch.h"

namespace pd {
namespace {

struct FitConst {
    double node[4];        // Chebyshev nodes cos(pi (i + 1/2) / 4)
    double vinv[4][4];     // inverse Vandermonde: monomial coefficient p = sum_i vinv[p][i] f(node_i)
    double chk[5];         // validation abscissae (none of them a node)
};

// exact thin-plate value of set `sid` at (M, a): the sum the step kernel evaluates (rows hold c/2,
// so a term is c r^2 log r^2), then the degree-1 tail with the stored shift and 1/scale
__device__ double tps_exact(const double *__restrict__ rows, const double2 *__restrict__ pts, int sid,
                            double M, double a) {
    const double *r = rows + (size_t)sid * 64;
    const unsigned char *ib = reinterpret_cast<const unsigned char *>(r + 57);
    double acc = 0.0;
    for (int k = 0; k < 50; ++k) {
        const double2 p = pts[ib[k]];
        const double dm = M - p.x, da = a - p.y;
        const double r2 = fma(dm, dm, da * da);
        if (r2 > 0.0) acc = fma(r[k] * r2, log(r2), acc);
    }
    const double xh = (M - r[53]) * r[55];
    const double yh = (a - r[54]) * r[56];
    return acc + r[50] + r[51] * xh + r[52] * yh;
}

__global__ void fit_kernel(PatchGridIn g, FitConst fc, const int2 *__restrict__ pbase, float *__restrict__ patch,
                           double tol, unsigned long long *n_failed, unsigned long long *max_err_bits) {
    const long long nsub = (long long)g.sub_x * g.sub_y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)g.nm * g.na * nsub) return;
    const int cell = (int)(t / nsub), sub = (int)(t % nsub);
    const int pb = pbase[cell].x;
    if (pb < 0) return;
    const int two = pb & 1;
    const int c = g.cells[cell];
    int sid[2];
    if (c >= 0) {
        sid[0] = c; sid[1] = c;
    } else {
        const int k = -c - 1;
        sid[0] = g.imp_id[k];
        sid[1] = (int)((g.imp_hint[k] >> 16) & 0xFFFF);
    }
    const int im = cell % g.nm, ia = cell / g.nm;
    const int sx = sub % g.sub_x, sy = sub / g.sub_x;
    const double hx = 0.5 * g.dm / g.sub_x, hy = 0.5 * g.da / g.sub_y;
    const double mc = g.m0 + g.dm * im + hx * (2 * sx + 1);
    const double ac = g.a0 + g.da * ia + hy * (2 * sy + 1);
    for (int w = 0; w <= two; ++w) {
        double F[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                F[i][j] = tps_exact(g.rows, g.points, sid[w], mc + hx * fc.node[i], ac + hy * fc.node[j]);
        // C[q][p]: coefficient of u^p v^q = sum_ij vinv[p][i] vinv[q][j] F[i][j]
        double T[4][4];
        for (int p = 0; p < 4; ++p)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int i = 0; i < 4; ++i) s = fma(fc.vinv[p][i], F[i][j], s);
                T[p][j] = s;
            }
        double C[4][4];
        for (int q = 0; q < 4; ++q)
            for (int p = 0; p < 4; ++p) {
                double s = 0.0;
                for (int j = 0; j < 4; ++j) s = fma(fc.vinv[q][j], T[p][j], s);
                C[q][p] = s;
            }
        // Stored form (64 B): 16 floats C[q][p]; the constant term keeps double precision as
        // hi = C[0][0], lo in the slot of the u^3 v^3 coefficient (1e-15 on these cells, dropped);
        // everything but the constant is the variation of the coefficient over a 2e-3-wide cell,
        // so the kernel evaluates it in fp32 (1e-10 absolute) and adds the constant in double.
        float S[16];
        for (int q = 0; q < 4; ++q)
            for (int p = 0; p < 4; ++p) S[q * 4 + p] = (float)C[q][p];
        S[15] = (float)(C[0][0] - (double)S[0]);
        double err = 0.0;
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                const float u = (float)fc.chk[i], v = (float)fc.chk[j];
                const float r0 = fmaf(fmaf(fmaf(S[3], u, S[2]), u, S[1]), u, 0.0f);
                const float r1 = fmaf(fmaf(fmaf(S[7], u, S[6]), u, S[5]), u, S[4]);
                const float r2 = fmaf(fmaf(fmaf(S[11], u, S[10]), u, S[9]), u, S[8]);
                const float r3 = fmaf(fmaf(S[14], u, S[13]), u, S[12]);
                const float var = fmaf(fmaf(fmaf(r3, v, r2), v, r1), v, r0);
                const double val = (double)S[0] + ((double)S[15] + (double)var);
                const double f = tps_exact(g.rows, g.points, sid[w], mc + hx * (double)u, ac + hy * (double)v);
                err = fmax(err, fabs(val - f));
            }
        float *out = patch + ((size_t)(pb >> 1) + (size_t)sub * (1 + two) + w) * 16;
        const bool bad = !(err <= tol);
        if (bad) S[0] = __int_as_float(0x7fc00000);
        for (int k = 0; k < 16; ++k) out[k] = S[k];
        if (bad) atomicAdd(n_failed, 1ULL);
        else atomicMax(max_err_bits, (unsigned long long)__double_as_longlong(err));
    }
}

}  // namespace

int build_patch_grid(const PatchGridIn &in, double tol, PatchGridOut *out, cudaStream_t st) {
    const size_t ncell = (size_t)in.nm * in.na;
    const long long nsub = (long long)in.sub_x * in.sub_y;
    std::vector<int2> pbase(ncell);
    long long n = 0;
    for (size_t c = 0; c < ncell; ++c) {
        const int v = in.cells_host[c];
        int sets = 1;
        if (v < 0) sets = (in.imp_hint_host[-v - 1] >> 63) ? 2 : 0;
        if (sets == 0) { pbase[c] = make_int2(-1, 0); continue; }
        if (n + nsub * sets >= (1LL << 30)) return 1;       // index packed into 31 bits with the flag
        pbase[c] = make_int2((int)((n << 1) | (sets - 1)),
                             sets == 2 ? (int)(in.imp_hint_host[-v - 1] & 0xFFFF) : 0);   // slots p << 8 | q
        n += nsub * sets;
    }
    out->n_patches = n;
    if (cudaMalloc(&out->pbase, ncell * sizeof(int2)) != cudaSuccess) return 1;
    if (cudaMalloc(&out->patch, (size_t)(n > 0 ? n : 1) * 16 * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(out->pbase, pbase.data(), ncell * sizeof(int2), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    unsigned long long *stats = nullptr;
    if (cudaMalloc(&stats, 2 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), st);
    FitConst fc;
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < 4; ++i) fc.node[i] = cos(PI * (i + 0.5) / 4.0);
    // Lagrange basis polynomials of the 4 nodes in monomial form = columns of the inverse Vandermonde
    for (int i = 0; i < 4; ++i) {
        double poly[4] = {1.0, 0.0, 0.0, 0.0};
        int deg = 0;
        double denom = 1.0;
        for (int k = 0; k < 4; ++k) {
            if (k == i) continue;
            denom *= fc.node[i] - fc.node[k];
            for (int p = deg + 1; p >= 1; --p) poly[p] = poly[p - 1] - fc.node[k] * poly[p];   // times (u - node_k)
            poly[0] = -fc.node[k] * poly[0];
            ++deg;
        }
        for (int p = 0; p < 4; ++p) fc.vinv[p][i] = poly[p] / denom;
    }
    const double chk[5] = {-1.0, -0.7, 0.0, 0.7, 1.0};
    for (int i = 0; i < 5; ++i) fc.chk[i] = chk[i];
    const long long threads = (long long)ncell * nsub;
    const int block = 128;
    fit_kernel<<<(unsigned)((threads + block - 1) / block), block, 0, st>>>(in, fc, out->pbase, out->patch, tol,
                                                                           stats, stats + 1);
    unsigned long long h[2] = {0, 0};
    if (cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    cudaFree(stats);
    out->n_failed = (long long)h[0];
    double me;
    memcpy(&me, &h[1], sizeof(me));
    out->max_err_kept = me;
    return cudaGetLastError() != cudaSuccess;
}

void free_patch_grid(PatchGridOut *p) {
    if (p->pbase) cudaFree(p->pbase);
    if (p->patch) cudaFree(p->patch);
    p->pbase = nullptr;
    p->patch = nullptr;
}

}  // namespace pd
