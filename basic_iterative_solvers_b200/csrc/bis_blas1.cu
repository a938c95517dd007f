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

// Input vectors of a streaming kernel.  The kernel reads ALL inputs of EW_UNROLL elements before it
// runs the per-element body (which stores): outputs may alias inputs element-wise
// (gauss_seidel.hpp:34, gmres.hpp:25, kernels.hpp:369), so the compiler cannot hoist a later element's
// loads above an earlier element's stores by itself, and a thread would have one element's loads in
// flight at a time.  Loading first gives NIN x EW_UNROLL independent loads per thread.
template <int NIN> struct EwIn {
    const double *p[NIN > 0 ? NIN : 1];
};

// Generic streaming kernel: prep() once per thread (loads the device scalars and forms
// alpha/beta/omega), then f(i, v, acc, scalars) per element with v[k] = in.p[k][i]; NRED fused reductions.
// Block b owns rows [b*chunk, (b+1)*chunk) -- a rule that depends on the global problem only, so the
// block partials of a fused reduction are the same numbers at every rank count (bis_internal.cuh).  A
// trip covers EW_THREADS * EW_UNROLL = 1024 consecutive rows; thread t takes rows t, t+256, t+512, t+768
// of the trip, in that order (coalesced 8-byte accesses: measured 6.2-6.3 TB/s, 0.95 of the copy peak).
template <int NRED, int NIN, class P, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(int64_t n, int chunk, P prep, EwIn<NIN> in, F f, RedArgs ra) {
    const auto sc = prep();
    double acc[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int q = 0; q < (NRED > 0 ? NRED : 1); ++q) acc[q] = 0.0;
    const int64_t base = (int64_t)blockIdx.x * chunk;
    const int64_t end = base + chunk < n ? base + chunk : n;
    constexpr int TRIP = EW_THREADS * EW_UNROLL;
    for (int64_t t0 = base; t0 < end; t0 += TRIP) {
        const int64_t i = t0 + threadIdx.x;
        if (t0 + TRIP <= end) {
            double v[EW_UNROLL][NIN > 0 ? NIN : 1];
#pragma unroll
            for (int u = 0; u < EW_UNROLL; ++u)
#pragma unroll
                for (int k = 0; k < NIN; ++k) v[u][k] = in.p[k][i + u * EW_THREADS];
#pragma unroll
            for (int u = 0; u < EW_UNROLL; ++u) f(i + u * EW_THREADS, v[u], acc, sc);
        } else {
#pragma unroll
            for (int u = 0; u < EW_UNROLL; ++u) {
                const int64_t e = i + u * EW_THREADS;
                if (e < end) {
                    double v[NIN > 0 ? NIN : 1];
#pragma unroll
                    for (int k = 0; k < NIN; ++k) v[k] = in.p[k][e];
                    f(e, v, acc, sc);
                }
            }
        }
    }
    if constexpr (NRED > 0) block_reduce_finish<NRED>(acc, ra);
}

struct Sc3 { double a, b, c; };

template <int NRED, int NIN, class P, class F>
int launch_ew2(bis_context *c, int64_t n, P prep, const EwIn<NIN> &in, F f, int slot_a = -1, int slot_b = -1) {
    RedArgs ra = bis_red_args(c, slot_a, slot_b);
    int chunk = 4096;   // plain streaming kernels: any rule does
    int64_t blocks = (n + chunk - 1) / chunk;
    if (NRED > 0) {
        // the launch shape of a reducing kernel is a function of the GLOBAL problem (bis_internal.cuh)
        const RowPartition part = bis_partition_for(c, n);
        chunk = part.chunk;
        while ((n + chunk - 1) / chunk > BIS_MAX_RED_BLOCKS) chunk <<= 1;
        blocks = (n + chunk - 1) / chunk;
        int off[BIS_NSLAB + 1];
        bool aligned = chunk == part.chunk;
        for (int i = 0; i <= part.n_slab; ++i) {
            off[i] = (int)((part.slab_row[i] + chunk - 1) / chunk);
            if (i < part.n_slab && part.slab_row[i] % chunk != 0) aligned = false;
        }
        RowPartition eff = part;
        if (!aligned) eff.invariant = false;
        bis_red_set_slabs(c, ra, eff, off, (int)(blocks > 0 ? blocks : 1));
    }
    if (blocks < 1) blocks = 1;
    BIS_CHECK(bis_prof_begin(c, BIS_PROF_VECTOR));
    ew_kernel<NRED, NIN, P, F><<<(unsigned)blocks, EW_THREADS, 0, c->stream>>>(n, chunk, prep, in, f, ra);
    BIS_LAUNCH_CHECK(c);
    BIS_CHECK(bis_prof_end(c, BIS_PROF_VECTOR));
    if (NRED > 0) BIS_CHECK(bis_reduce_finish(c, slot_a, slot_b));
    return 0;
}

template <int NRED, int NIN, class F>
int launch_ew(bis_context *c, int64_t n, const EwIn<NIN> &in, F f, int slot_a = -1, int slot_b = -1) {
    return launch_ew2<NRED, NIN>(c, n, [] __device__() { return 0; }, in,
                                 [=] __device__(int64_t i, const double *v, double *acc, int) { f(i, v, acc); },
                                 slot_a, slot_b);
}

inline bool slot_ok(int s) { return s >= 0 && s < BIS_NUM_SCALARS; }

} // namespace

#define REQ_CTX(c) BIS_REQUIRE((c) != nullptr, "null context")
#define REQ_SLOT(s) BIS_REQUIRE(slot_ok(s), "bad scalar slot %d", (s))

// ---- kernels.hpp:119-153 ------------------------------------------------------
extern "C" int bis_subtract_vectors(bis_context *c, double *out, const double *a, const double *b,
                                    int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0, 2>(c, n, EwIn<2>{{a, b}},
                           [=] __device__(int64_t i, const double *v, double *) { out[i] = fma(-scale, v[1], v[0]); });
}
extern "C" int bis_sum_vectors(bis_context *c, double *out, const double *a, const double *b,
                               int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0, 2>(c, n, EwIn<2>{{a, b}},
                           [=] __device__(int64_t i, const double *v, double *) { out[i] = fma(scale, v[1], v[0]); });
}
extern "C" int bis_elemwise_mult_vectors(bis_context *c, double *out, const double *a,
                                         const double *b, int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0, 2>(c, n, EwIn<2>{{a, b}}, [=] __device__(int64_t i, const double *v, double *) {
        out[i] = mul_rn(mul_rn(v[0], scale), v[1]);
    });
}
extern "C" int bis_elemwise_div_vectors(bis_context *c, double *out, const double *a,
                                        const double *b, int64_t n, double scale) {
    REQ_CTX(c);
    return launch_ew<0, 2>(c, n, EwIn<2>{{a, b}}, [=] __device__(int64_t i, const double *v, double *) {
        out[i] = div_rn(v[0], mul_rn(scale, v[1]));
    });
}
// kernels.hpp:214-220
extern "C" int bis_scale(bis_context *c, double *out, const double *v_in, double scalar, int64_t n) {
    REQ_CTX(c);
    return launch_ew<0, 1>(c, n, EwIn<1>{{v_in}},
                           [=] __device__(int64_t i, const double *v, double *) { out[i] = mul_rn(v[0], scalar); });
}
// kernels.hpp:236-241
extern "C" int bis_init_vector(bis_context *c, double *v_out, double value, int64_t n) {
    REQ_CTX(c);
    return launch_ew<0, 0>(c, n, EwIn<0>{{nullptr}},
                           [=] __device__(int64_t i, const double *, double *) { v_out[i] = value; });
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
    return launch_ew<0, 4>(c, n, EwIn<4>{{x_new, x_old, D, b}}, [=] __device__(int64_t i, const double *v, double *) {
        double scaled = mul_rn(v[2], v[1]);
        double adj = sub_rn(v[0], scaled);
        x_new[i] = div_rn(sub_rn(v[3], adj), v[2]);
    });
}

// ---- reductions: kernels.hpp:194-212 ------------------------------------------
extern "C" int bis_dot_to_slot(bis_context *c, const double *a, const double *b, int64_t n, int slot) {
    REQ_CTX(c);
    REQ_SLOT(slot);
    return launch_ew<1, 2>(c, n, EwIn<2>{{a, b}},
                           [=] __device__(int64_t, const double *v, double *acc) { acc[0] = fma(v[0], v[1], acc[0]); },
                           slot);
}
extern "C" int bis_sumsq_to_slot(bis_context *c, const double *v_in, int64_t n, int slot) {
    REQ_CTX(c);
    REQ_SLOT(slot);
    return launch_ew<1, 1>(c, n, EwIn<1>{{v_in}},
                           [=] __device__(int64_t, const double *v, double *acc) { acc[0] = fma(v[0], v[0], acc[0]); },
                           slot);
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
    // x_new == NULL: the x update is left to bis_cg_direction_x, which reads p_old anyway (one pass
    // over p_old less per iteration); the inputs x_old / p_old are then not touched here
    const bool with_x = x_new != nullptr;
    const double *xo = with_x ? x_old : r_old, *po = with_x ? p_old : r_old;
    const EwIn<4> in{{xo, po, r_old, Ap}};
    if (precond == BIS_PRECOND_NONE) {
        REQ_SLOT(slot_rz_new);
        // z_new = r_new (copy_vector, kernels.hpp:396-399); (r,z) == (r,r): the same sum lands in both slots
        auto body = [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
            if (with_x) x_new[i] = fma(alpha, v[1], v[0]);
            double r = fma(-alpha, v[3], v[2]);
            r_new[i] = r;
            z_new[i] = r;
            acc[0] = fma(r, r, acc[0]);
            acc[1] = acc[0];
        };
        if (with_x) return launch_ew2<2, 4>(c, n, prep, in, body, slot_rr, slot_rz_new);
        return launch_ew2<2, 2>(c, n, prep, EwIn<2>{{r_old, Ap}},
                                [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
            double r = fma(-alpha, v[1], v[0]);
            r_new[i] = r;
            z_new[i] = r;
            acc[0] = fma(r, r, acc[0]);
            acc[1] = acc[0];
        }, slot_rr, slot_rz_new);
    }
    if (precond == BIS_PRECOND_JACOBI) {
        REQ_SLOT(slot_rz_new);
        if (with_x)
            return launch_ew2<2, 5>(c, n, prep, EwIn<5>{{x_old, p_old, r_old, Ap, A_D}},
                                    [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
                x_new[i] = fma(alpha, v[1], v[0]);
                double r = fma(-alpha, v[3], v[2]);
                r_new[i] = r;
                double z = div_rn(r, v[4]);   // elemwise_div_vectors with scale = 1.0 (1.0*d == d)
                z_new[i] = z;
                acc[0] = fma(r, r, acc[0]);
                acc[1] = fma(r, z, acc[1]);
            }, slot_rr, slot_rz_new);
        return launch_ew2<2, 3>(c, n, prep, EwIn<3>{{r_old, Ap, A_D}},
                                [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
            double r = fma(-alpha, v[1], v[0]);
            r_new[i] = r;
            double z = div_rn(r, v[2]);
            z_new[i] = z;
            acc[0] = fma(r, r, acc[0]);
            acc[1] = fma(r, z, acc[1]);
        }, slot_rr, slot_rz_new);
    }
    if (with_x)
        return launch_ew2<1, 4>(c, n, prep, in, [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
            x_new[i] = fma(alpha, v[1], v[0]);
            double r = fma(-alpha, v[3], v[2]);
            r_new[i] = r;
            acc[0] = fma(r, r, acc[0]);
        }, slot_rr);
    return launch_ew2<1, 2>(c, n, prep, EwIn<2>{{r_old, Ap}},
                            [=] __device__(int64_t i, const double *v, double *acc, double alpha) {
        double r = fma(-alpha, v[1], v[0]);
        r_new[i] = r;
        acc[0] = fma(r, r, acc[0]);
    }, slot_rr);
}

extern "C" int bis_cg_direction(bis_context *c, int64_t n, double *p_new, const double *z_new,
                                const double *p_old, int slot_rz_new, int slot_rz) {
    REQ_CTX(c);
    REQ_SLOT(slot_rz_new); REQ_SLOT(slot_rz);
    const double *S = c->d_scalars;
    return launch_ew2<0, 2>(c, n, [=] __device__() { return div_rn(S[slot_rz_new], S[slot_rz]); },
                            EwIn<2>{{z_new, p_old}},
                            [=] __device__(int64_t i, const double *v, double *, double beta) {
                                p_new[i] = fma(beta, v[1], v[0]);
                            });
}

// cg.hpp:27-28 and :47-52 in one pass over p_old: x_new = x_old + alpha p_old ; p_new = z_new + beta p_old
extern "C" int bis_cg_direction_x(bis_context *c, int64_t n, double *p_new, const double *z_new,
                                  const double *p_old, double *x_new, const double *x_old,
                                  int slot_rz_new, int slot_rz, int slot_pAp) {
    REQ_CTX(c);
    REQ_SLOT(slot_rz_new); REQ_SLOT(slot_rz); REQ_SLOT(slot_pAp);
    BIS_REQUIRE(x_new && x_old, "bis_cg_direction_x: null x");
    const double *S = c->d_scalars;
    return launch_ew2<0, 3>(c, n, [=] __device__() {
        Sc3 sc;
        sc.a = div_rn(S[slot_rz], S[slot_pAp]);       // alpha of this iteration (cg.hpp:19-23)
        sc.b = div_rn(S[slot_rz_new], S[slot_rz]);    // beta (cg.hpp:47)
        sc.c = 0.0;
        return sc;
    }, EwIn<3>{{z_new, p_old, x_old}}, [=] __device__(int64_t i, const double *v, double *, Sc3 sc) {
        x_new[i] = fma(sc.a, v[1], v[2]);
        p_new[i] = fma(sc.b, v[1], v[0]);
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
    return launch_ew2<0, 3>(c, n, [=] __device__() { return div_rn(S[slot_rho_old], S[slot_r0v]); },
                            EwIn<3>{{r_old, v, mode == 1 ? A_D : r_old}},
                            [=] __device__(int64_t i, const double *w, double *, double alpha) {
                                double sv = fma(-alpha, w[1], w[0]);
                                s[i] = sv;
                                if (mode == 0) s_tmp[i] = sv;
                                else if (mode == 1) s_tmp[i] = div_rn(sv, w[2]);
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
    return launch_ew2<2, 6>(c, n, [=] __device__() {
        Sc3 sc;
        sc.a = div_rn(S[slot_rho_old], S[slot_r0v]);   // alpha, bicgstab.hpp:34
        sc.b = div_rn(S[slot_zs], S[slot_zz]);         // omega, bicgstab.hpp:51
        sc.c = 0.0;
        return sc;
    }, EwIn<6>{{x_old, y, s_tmp, s, z, r0}}, [=] __device__(int64_t i, const double *v, double *acc, Sc3 sc) {
        double hv = fma(sc.a, v[1], v[0]);
        if (h) h[i] = hv;
        x_new[i] = fma(sc.b, v[2], hv);
        double r = fma(-sc.b, v[4], v[3]);
        r_new[i] = r;
        acc[0] = fma(v[5], r, acc[0]);
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
    return launch_ew2<0, 4>(c, n, [=] __device__() {
        Sc3 sc;
        double alpha = div_rn(S[slot_rho_old], S[slot_r0v]);
        sc.b = div_rn(S[slot_zs], S[slot_zz]);   // omega
        // beta = (rho_new / rho_old) * (alpha / omega), bicgstab.hpp:70
        sc.a = mul_rn(div_rn(S[slot_rho_new], S[slot_rho_old]), div_rn(alpha, sc.b));
        sc.c = 0.0;
        return sc;
    }, EwIn<4>{{p_old, v, r_new, mode == 1 ? A_D : r_new}}, [=] __device__(int64_t i, const double *w, double *, Sc3 sc) {
        double t = fma(-sc.b, w[1], w[0]);
        if (tmp) tmp[i] = t;
        double p = fma(sc.a, t, w[2]);
        p_new[i] = p;
        if (mode == 0) y_next[i] = p;
        else if (mode == 1) y_next[i] = div_rn(p, w[3]);
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
        return launch_ew2<1, 3>(c, n, prep, EwIn<3>{{w, v_j, v_next}},
                                [=] __device__(int64_t i, const double *v, double *acc, double h) {
            double t = fma(-h, v[1], v[0]);
            w[i] = t;
            acc[0] = fma(t, v[2], acc[0]);
        }, slot_out);
    return launch_ew2<1, 2>(c, n, prep, EwIn<2>{{w, v_j}},
                            [=] __device__(int64_t i, const double *v, double *acc, double h) {
        double t = fma(-h, v[1], v[0]);
        w[i] = t;
        acc[0] = fma(t, t, acc[0]);
    }, slot_out);
}

extern "C" int bis_scale_inv_norm(bis_context *c, int64_t n, double *out, const double *w,
                                  int slot_sumsq) {
    REQ_CTX(c);
    REQ_SLOT(slot_sumsq);
    const double *S = c->d_scalars;
    return launch_ew2<0, 1>(c, n, [=] __device__() { return div_rn(1.0, sqrt(S[slot_sumsq])); }, EwIn<1>{{w}},
                            [=] __device__(int64_t i, const double *v, double *, double inv) { out[i] = mul_rn(v[0], inv); });
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
    return launch_ew<0, 1>(c, n, EwIn<1>{{x_old}}, [=] __device__(int64_t i, const double *v, double *) {
        // dgemm_transpose1 (kernels.hpp:259-271): left-to-right, unfused
        double t = 0.0;
        for (int j = 0; j < k; ++j) t = add_rn(t, mul_rn(V[(int64_t)j * n + i], yc.y[j]));
        if (Vy) Vy[i] = t;
        x[i] = add_rn(v[0], t);
    });
}
