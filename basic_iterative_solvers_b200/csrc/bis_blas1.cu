// bis_blas1.cu -- BLAS-1 family (kernels.hpp:119-153, 194-257), normalize_x
// (methods/jacobi.hpp:27-40) and the fused per-method vector kernels of
// methods/{cg,bicgstab,gmres}.hpp.  All kernels are HBM-bound streaming
// kernels: 8 B per vector element read or written, no reuse.
//
// Rounding is pinned (library built with --fmad=false): the reference's
// sum/subtract_vectors are one FMA, everything else is separately rounded
// (oracle/port/bis_oracle.c header).
#include "bis_device.cuh"

namespace {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

// Generic streaming kernel: P() once per thread (loads the device scalars and
// forms alpha/beta/omega), then F(i, acc, scalars) per element; NRED fused
// reductions.
template <int NRED, class P, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(int64_t n, P prep, F f, RedArgs ra) {
    const auto sc = prep();
    double acc[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int q = 0; q < (NRED > 0 ? NRED : 1); ++q) acc[q] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * EW_THREADS;
    int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x;
    // main body: EW_UNROLL independent elements per thread per trip
    for (; i + (EW_UNROLL - 1) * stride < n; i += EW_UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) f(i + u * stride, acc, sc);
    }
    for (; i < n; i += stride) f(i, acc, sc);
    if constexpr (NRED > 0) block_reduce_finish<NRED>(acc, ra);
}

struct Sc3 { double a, b, c; };

template <int NRED, class P, class F>
int launch_ew2(bis_context *c, int64_t n, P prep, F f, int slot_a = -1, int slot_b = -1) {
    // fixed launch shape for a given n => bit-reproducible reductions
    int cap = c->sm_count * 8;
    if (cap > BIS_MAX_RED_BLOCKS) cap = BIS_MAX_RED_BLOCKS;
    int blocks = bis_blocks_for(n, EW_THREADS * EW_UNROLL, cap);
    RedArgs ra = bis_red_args(c, slot_a, slot_b);
    ra.total_blocks = blocks;
    BIS_CHECK(bis_prof_begin(c, BIS_PROF_VECTOR));
    ew_kernel<NRED, P, F><<<blocks, EW_THREADS, 0, c->stream>>>(n, prep, f, ra);
    BIS_LAUNCH_CHECK(c);
    BIS_CHECK(bis_prof_end(c, BIS_PROF_VECTOR));
    if (NRED > 0) BIS_CHECK(bis_reduce_finish(c, slot_a, slot_b));
    return 0;
}

template <int NRED, class F>
int launch_ew(bis_context *c, int64_t n, F f, int slot_a = -1, int slot_b = -1) {
    return launch_ew2<NRED>(c, n, [] __device__() { return 0; },
                            [=] __device__(int64_t i, double *acc, int) { f(i, acc); }, slot_a, slot_b);
}

inline bool slot_ok(int s) { return s >= 0 && s < BIS_NUM_SCALARS; }

} // namespace

#define REQ_CTX(c) BIS_REQUIRE((c) != nullptr, "null context")
#define REQ_SLOT(s) BIS_REQUIRE(slot_ok(s), "bad scalar slot %d", (s))

// ---- kernels.hpp:119-153 ------------------------------------------------------
extern "C" int bis_subtract_vectors(bis_context *c, double *out, const double *a, const double *b,
                                    int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) { out[i] = fma(-scale, b[i], a[i]); });
}
extern "C" int bis_sum_vectors(bis_context *c, double *out, const double *a, const double *b,
                               int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) { out[i] = fma(scale, b[i], a[i]); });
}
extern "C" int bis_elemwise_mult_vectors(bis_context *c, double *out, const double *a,
                                         const double *b, int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) {
        out[i] = mul_rn(mul_rn(a[i], scale), b[i]);
    });
}
extern "C" int bis_elemwise_div_vectors(bis_context *c, double *out, const double *a,
                                        const double *b, int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) {
        out[i] = div_rn(a[i], mul_rn(scale, b[i]));
    });
}
// kernels.hpp:214-220
extern "C" int bis_scale(bis_context *c, double *out, const double *v, double scalar, int64_t n) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) { out[i] = mul_rn(v[i], scalar); });
}
// kernels.hpp:236-241
extern "C" int bis_init_vector(bis_context *c, double *v, double value, int64_t n) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) { v[i] = value; });
}
// kernels.hpp:252-257
extern "C" int bis_copy_vector(bis_context *c, double *out, const double *in, int64_t n) {
    REQ_CTX(c);
    if (out == in || n == 0) return 0;
    BIS_CUDA(cudaMemcpyAsync(out, in, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
// methods/jacobi.hpp:27-40
extern "C" int bis_normalize_x(bis_context *c, double *x_new, const double *x_old, const double *D,
                               const double *b, int64_t n) {
    REQ_CTX(c);
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) {
        double d = D[i];
        double scaled = mul_rn(d, x_old[i]);
        double adj = sub_rn(x_new[i], scaled);
        x_new[i] = div_rn(sub_rn(b[i], adj), d);
    });
}

// ---- reductions: kernels.hpp:194-212 ------------------------------------------
extern "C" int bis_dot_to_slot(bis_context *c, const double *a, const double *b, int64_t n, int slot) {
    REQ_CTX(c);
    REQ_SLOT(slot);
    return launch_ew<1>(c, n, [=] __device__(int64_t i, double *acc) { acc[0] = fma(a[i], b[i], acc[0]); },
                        slot);
}
extern "C" int bis_sumsq_to_slot(bis_context *c, const double *v, int64_t n, int slot) {
    REQ_CTX(c);
    REQ_SLOT(slot);
    return launch_ew<1>(c, n, [=] __device__(int64_t i, double *acc) {
        double t = v[i];
        acc[0] = fma(t, t, acc[0]);
    }, slot);
}
extern "C" int bis_dot(bis_context *c, const double *a, const double *b, int64_t n, double *result) {
    REQ_CTX(c);
    BIS_REQUIRE(result, "bis_dot: null result");
    BIS_CHECK(bis_dot_to_slot(c, a, b, n, BIS_NUM_SCALARS - 1));
    return bis_scalar_get(c, BIS_NUM_SCALARS - 1, 1, result);
}
extern "C" int bis_euclidean_vec_norm(bis_context *c, const double *v, int64_t n, double *result) {
    REQ_CTX(c);
    BIS_REQUIRE(result, "bis_euclidean_vec_norm: null result");
    BIS_CHECK(bis_sumsq_to_slot(c, v, n, BIS_NUM_SCALARS - 1));
    BIS_CHECK(bis_scalar_get(c, BIS_NUM_SCALARS - 1, 1, result));
    *result = sqrt(*result);
    return 0;
}

// ---- CG: methods/cg.hpp:6-54 ---------------------------------------------------
extern "C" int bis_cg_update(bis_context *c, int precond, int64_t n, double *x_new,
                             const double *x_old, const double *p_old, double *r_new,
                             const double *r_old, const double *Ap, double *z_new,
                             const double *A_D, int slot_rz, int slot_pAp, int slot_rr,
                             int slot_rz_new) {
    REQ_CTX(c);
    REQ_SLOT(slot_rz); REQ_SLOT(slot_pAp); REQ_SLOT(slot_rr);
    const double *S = c->d_scalars;
    // alpha <- (r_old, z_old) / (Ap_old, p_old), cg.hpp:19-23
    auto prep = [=] __device__() { return div_rn(S[slot_rz], S[slot_pAp]); };
    if (precond == BIS_PRECOND_NONE) {
        REQ_SLOT(slot_rz_new);
        // z_new = r_new (copy_vector, kernels.hpp:396-399); (r,z) == (r,r)
        BIS_CHECK((launch_ew2<1>(c, n, prep, [=] __device__(int64_t i, double *acc, double alpha) {
            x_new[i] = fma(alpha, p_old[i], x_old[i]);
            double r = fma(-alpha, Ap[i], r_old[i]);
            r_new[i] = r;
            z_new[i] = r;
            acc[0] = fma(r, r, acc[0]);
        }, slot_rr)));
        return bis_scalar_copy(c, slot_rz_new, slot_rr);
    }
    if (precond == BIS_PRECOND_JACOBI) {
        REQ_SLOT(slot_rz_new);
        return launch_ew2<2>(c, n, prep, [=] __device__(int64_t i, double *acc, double alpha) {
            x_new[i] = fma(alpha, p_old[i], x_old[i]);
            double r = fma(-alpha, Ap[i], r_old[i]);
            r_new[i] = r;
            double z = div_rn(r, A_D[i]);   // elemwise_div_vectors with scale = 1.0 (1.0*d == d)
            z_new[i] = z;
            acc[0] = fma(r, r, acc[0]);
            acc[1] = fma(r, z, acc[1]);
        }, slot_rr, slot_rz_new);
    }
    return launch_ew2<1>(c, n, prep, [=] __device__(int64_t i, double *acc, double alpha) {
        x_new[i] = fma(alpha, p_old[i], x_old[i]);
        double r = fma(-alpha, Ap[i], r_old[i]);
        r_new[i] = r;
        acc[0] = fma(r, r, acc[0]);
    }, slot_rr);
}

extern "C" int bis_cg_direction(bis_context *c, int64_t n, double *p_new, const double *z_new,
                                const double *p_old, int slot_rz_new, int slot_rz) {
    REQ_CTX(c);
    REQ_SLOT(slot_rz_new); REQ_SLOT(slot_rz);
    const double *S = c->d_scalars;
    return launch_ew2<0>(c, n, [=] __device__() { return div_rn(S[slot_rz_new], S[slot_rz]); },
                         [=] __device__(int64_t i, double *, double beta) {
                             p_new[i] = fma(beta, p_old[i], z_new[i]);
                         });
}

// ---- BiCGSTAB: methods/bicgstab.hpp:8-83 ----------------------------------------
extern "C" int bis_bicgstab_s(bis_context *c, int precond, int64_t n, double *s, double *s_tmp,
                              const double *r_old, const double *v, const double *A_D,
                              int slot_rho_old, int slot_r0v) {
    REQ_CTX(c);
    REQ_SLOT(slot_rho_old); REQ_SLOT(slot_r0v);
    const double *S = c->d_scalars;
    const int mode = (precond == BIS_PRECOND_NONE) ? 0 : (precond == BIS_PRECOND_JACOBI ? 1 : 2);
    return launch_ew2<0>(c, n, [=] __device__() { return div_rn(S[slot_rho_old], S[slot_r0v]); },
                         [=] __device__(int64_t i, double *, double alpha) {
                             double sv = fma(-alpha, v[i], r_old[i]);
                             s[i] = sv;
                             if (mode == 0) s_tmp[i] = sv;
                             else if (mode == 1) s_tmp[i] = div_rn(sv, A_D[i]);
                         });
}

extern "C" int bis_bicgstab_xr(bis_context *c, int64_t n, double *h, double *x_new,
                               const double *x_old, const double *y, const double *s_tmp,
                               double *r_new, const double *s, const double *z, const double *r0,
                               int slot_rho_old, int slot_r0v, int slot_zs, int slot_zz,
                               int slot_rho_new, int slot_rr) {
    REQ_CTX(c);
    REQ_SLOT(slot_rho_old); REQ_SLOT(slot_r0v); REQ_SLOT(slot_zs); REQ_SLOT(slot_zz);
    REQ_SLOT(slot_rho_new); REQ_SLOT(slot_rr);
    const double *S = c->d_scalars;
    return launch_ew2<2>(c, n, [=] __device__() {
        Sc3 sc;
        sc.a = div_rn(S[slot_rho_old], S[slot_r0v]);   // alpha, bicgstab.hpp:34
        sc.b = div_rn(S[slot_zs], S[slot_zz]);         // omega, bicgstab.hpp:51
        sc.c = 0.0;
        return sc;
    }, [=] __device__(int64_t i, double *acc, Sc3 sc) {
        double hv = fma(sc.a, y[i], x_old[i]);
        if (h) h[i] = hv;
        x_new[i] = fma(sc.b, s_tmp[i], hv);
        double r = fma(-sc.b, z[i], s[i]);
        r_new[i] = r;
        acc[0] = fma(r0[i], r, acc[0]);
        acc[1] = fma(r, r, acc[1]);
    }, slot_rho_new, slot_rr);
}

extern "C" int bis_bicgstab_p(bis_context *c, int precond, int64_t n, double *tmp, double *p_new,
                              const double *p_old, const double *v, const double *r_new,
                              double *y_next, const double *A_D, int slot_rho_new,
                              int slot_rho_old, int slot_r0v, int slot_zs, int slot_zz) {
    REQ_CTX(c);
    REQ_SLOT(slot_rho_new); REQ_SLOT(slot_rho_old); REQ_SLOT(slot_r0v); REQ_SLOT(slot_zs);
    REQ_SLOT(slot_zz);
    const double *S = c->d_scalars;
    const int mode = !y_next ? 2 : ((precond == BIS_PRECOND_NONE) ? 0 : (precond == BIS_PRECOND_JACOBI ? 1 : 2));
    return launch_ew2<0>(c, n, [=] __device__() {
        Sc3 sc;
        double alpha = div_rn(S[slot_rho_old], S[slot_r0v]);
        sc.b = div_rn(S[slot_zs], S[slot_zz]);   // omega
        // beta = (rho_new / rho_old) * (alpha / omega), bicgstab.hpp:70
        sc.a = mul_rn(div_rn(S[slot_rho_new], S[slot_rho_old]), div_rn(alpha, sc.b));
        sc.c = 0.0;
        return sc;
    }, [=] __device__(int64_t i, double *, Sc3 sc) {
        double t = fma(-sc.b, v[i], p_old[i]);
        if (tmp) tmp[i] = t;
        double p = fma(sc.a, t, r_new[i]);
        p_new[i] = p;
        if (mode == 0) y_next[i] = p;
        else if (mode == 1) y_next[i] = div_rn(p, A_D[i]);
    });
}

// ---- GMRES: methods/gmres.hpp ----------------------------------------------------
extern "C" int bis_mgs_step(bis_context *c, int64_t n, double *w, const double *v_j,
                            const double *v_next, int slot_h_j, int slot_out) {
    REQ_CTX(c);
    REQ_SLOT(slot_h_j); REQ_SLOT(slot_out);
    const double *S = c->d_scalars;
    auto prep = [=] __device__() { return S[slot_h_j]; };
    if (v_next)
        return launch_ew2<1>(c, n, prep, [=] __device__(int64_t i, double *acc, double h) {
            double t = fma(-h, v_j[i], w[i]);
            w[i] = t;
            acc[0] = fma(t, v_next[i], acc[0]);
        }, slot_out);
    return launch_ew2<1>(c, n, prep, [=] __device__(int64_t i, double *acc, double h) {
        double t = fma(-h, v_j[i], w[i]);
        w[i] = t;
        acc[0] = fma(t, t, acc[0]);
    }, slot_out);
}

extern "C" int bis_scale_inv_norm(bis_context *c, int64_t n, double *out, const double *w,
                                  int slot_sumsq) {
    REQ_CTX(c);
    REQ_SLOT(slot_sumsq);
    const double *S = c->d_scalars;
    return launch_ew2<0>(c, n, [=] __device__() { return div_rn(1.0, sqrt(S[slot_sumsq])); },
                         [=] __device__(int64_t i, double *, double inv) { out[i] = mul_rn(w[i], inv); });
}

namespace {
struct YCoef { double y[64]; };
}

extern "C" int bis_gmres_update_x(bis_context *c, int64_t n, int k, const double *V,
                                  const double *y, double *x, const double *x_old, double *Vy) {
    REQ_CTX(c);
    BIS_REQUIRE(k >= 0 && k <= 64, "bis_gmres_update_x: k=%d outside [0,64]", k);
    BIS_REQUIRE(k == 0 || y, "bis_gmres_update_x: null y");
    YCoef yc;
    for (int j = 0; j < 64; ++j) yc.y[j] = (j < k) ? y[j] : 0.0;
    return launch_ew<0>(c, n, [=] __device__(int64_t i, double *) {
        // dgemm_transpose1 (kernels.hpp:259-271): left-to-right, unfused
        double t = 0.0;
        for (int j = 0; j < k; ++j) t = add_rn(t, mul_rn(V[(int64_t)j * n + i], yc.y[j]));
        if (Vy) Vy[i] = t;
        x[i] = add_rn(x_old[i], t);
    });
}
