// compress_plan.h -- eigenvalue-space compression of the per-SNP Gram rows (host code, plain C++).
//
// Every SNP-dependent number a likelihood evaluation needs is a weighted moment
//     f_p(lambda) = sum_l a_l / (lambda d_l + 1)^p ,   a_l = x_l w_jl   or   a_l = x_l^2 ,   p = 1, 2, 3
// over the n eigenvalues d_l (reference: the three dsyrk calls of precompute_mat, pygemma_model.pyx:938,
// :943, :1002, which redo these sums for every SNP and every lambda).  g(u) = (lambda e^u + 1)^-p is
// analytic in the strip |Im u| < pi for every lambda > 0, so on a short interval of u = log d it is
// reproduced to rounding error by a degree-(q-1) Chebyshev interpolant whose error does not depend on
// lambda:  g(u_l) = sum_k L_k(u_l) g(u_k).  Summing over the eigenvalues of the interval first,
//     f_p(lambda) = sum_k z_k / (lambda d_k + 1)^p ,   z_k = sum_l L_k(log d_l) a_l ,
// i.e. the n terms collapse onto q nodes per interval once per SNP, independent of lambda and p, and every
// optimiser iteration afterwards costs O(#nodes) instead of O(n).  With kCq = 10 nodes on intervals of an
// eighth of a decade the Bernstein-ellipse parameter is rho ~ 2 pi / (ln(10)/16) ~ 44; measured worst-case
// error over lambda in [1e-5, 1e5], relative to sum_l |a_l| h_l^p (tests/test_host_logic.py):
// 1e-17 (p = 1), 2e-16 (p = 2), 1.4e-14 (p = 3, the pole is of third order; p = 3 feeds only Newton's second
// derivative) -- at or below the rounding error of the direct n-term sum.
//
// Plan (all index ranges refer to eigenvalues sorted ascending):
//   * eigenvalues with lambda_max * d <= 2^-54 give 1/(lambda d + 1) == 1.0 in fp64 for every lambda the
//     optimiser can visit (lambda <= 1e5): they collapse exactly onto one node d = 0;
//   * a maximal run of eigenvalues inside [d_i, d_i * 10^(1/kCSub)] with more than kCq members becomes a
//     COMPRESS segment with kCq Chebyshev nodes on [log d_min, log d_max] of the run (one node if all equal);
//   * everything else is kept exactly: consecutive leftovers form COPY segments, one node per eigenvalue.
#pragma once

#include <math.h>

#include <algorithm>
#include <vector>

namespace pg {

constexpr int kCq = 10;            // Chebyshev nodes per compressed segment
constexpr int kCSub = 8;           // segments per decade of eigenvalue
constexpr double kLambdaMax = 1e5; // largest lambda the optimiser evaluates (pyx:146, :158)

enum SegType { kSegCopy = 0, kSegCompress = 1 };

struct Segment {
    int l0, l1;  // sorted-eigenvalue index range [l0, l1)
    int kb;      // first node
    int kq;      // nodes: l1-l0 for COPY; 1 or kCq for COMPRESS
    int type;
};

struct CompressPlan {
    int n = 0, Kc = 0;
    std::vector<double> nodes;   // [Kc] node eigenvalues
    std::vector<Segment> segs;
    std::vector<double> Lw;      // [n][kCq] interpolation weights (COMPRESS rows; unused entries 0)
    std::vector<int> seg_of;     // [n]
};

inline double plan_cheb_node(int k) { return cos(3.14159265358979323846 * (k + 0.5) / kCq); }

// L_k(x), k = 0..kCq-1, for first-kind Chebyshev nodes: L_k(x) = (1/q) sum_j eps_j T_j(x_k) T_j(x)
inline void plan_lagrange(double x, double* L)
{
    long double T[kCq];
    T[0] = 1.0L; T[1] = x;
    for (int j = 2; j < kCq; ++j) T[j] = 2.0L * x * T[j - 1] - T[j - 2];
    for (int k = 0; k < kCq; ++k) {
        const long double th = 3.14159265358979323846264338327950288L * (k + 0.5L) / kCq;
        long double s = T[0];
        for (int j = 1; j < kCq; ++j) s += 2.0L * cosl(j * th) * T[j];
        L[k] = (double)(s / kCq);
    }
}

// d: n eigenvalues, ascending, >= 0
inline void build_compress_plan(const double* d, int n, CompressPlan* P)
{
    P->n = n; P->Kc = 0;
    P->nodes.clear(); P->segs.clear();
    P->Lw.assign((size_t)n * kCq, 0.0);
    P->seg_of.assign(n, -1);
    const double tiny = ldexp(1.0, -54) / kLambdaMax;
    const double ratio = pow(10.0, 1.0 / kCSub);
    auto add_copy = [&](int l) {
        if (!P->segs.empty() && P->segs.back().type == kSegCopy && P->segs.back().l1 == l) {
            P->segs.back().l1 = l + 1;
            P->segs.back().kq++;
        } else {
            P->segs.push_back(Segment{l, l + 1, (int)P->nodes.size(), 1, kSegCopy});
        }
        P->seg_of[l] = (int)P->segs.size() - 1;
        P->nodes.push_back(d[l]);
    };
    int i = 0;
    while (i < n && d[i] <= tiny) ++i;
    if (i > 0) {
        P->segs.push_back(Segment{0, i, 0, 1, kSegCompress});
        P->nodes.push_back(0.0);
        for (int l = 0; l < i; ++l) { P->Lw[(size_t)l * kCq] = 1.0; P->seg_of[l] = 0; }
    }
    while (i < n) {
        const double hi = d[i] * ratio;
        int j = i;
        while (j < n && d[j] <= hi) ++j;
        const int cnt = j - i;
        if (cnt <= kCq) {
            for (int l = i; l < j; ++l) add_copy(l);
        } else if (d[j - 1] == d[i]) {
            P->segs.push_back(Segment{i, j, (int)P->nodes.size(), 1, kSegCompress});
            P->nodes.push_back(d[i]);
            for (int l = i; l < j; ++l) { P->Lw[(size_t)l * kCq] = 1.0; P->seg_of[l] = (int)P->segs.size() - 1; }
        } else {
            const double ua = log(d[i]), ub = log(d[j - 1]);
            P->segs.push_back(Segment{i, j, (int)P->nodes.size(), kCq, kSegCompress});
            for (int k = 0; k < kCq; ++k) P->nodes.push_back(exp(0.5 * (ua + ub) + 0.5 * (ub - ua) * plan_cheb_node(k)));
            for (int l = i; l < j; ++l) {
                double x = (2.0 * log(d[l]) - ua - ub) / (ub - ua);
                x = std::min(1.0, std::max(-1.0, x));
                plan_lagrange(x, &P->Lw[(size_t)l * kCq]);
                P->seg_of[l] = (int)P->segs.size() - 1;
            }
        }
        i = j;
    }
    P->Kc = (int)P->nodes.size();
}

// ---- moments straight from the fused rotation (rotate_i8_tc2.cuh, FUSE) ----------------------------------------------
// The epilogue of an eigen tile squares the rotated values and accumulates, per thread, the x^2 moments of the
// `piece_eig` eigenvectors it holds: one partial sum ("piece") per (piece_eig-aligned chunk, COMPRESS segment).  A segment's
// node k is the sum of its pieces in eigen order (moments_reduce_kernel); COPY rows go straight to their node.
struct FusedRow {
    int id;   // COMPRESS row: piece index; COPY row: node
    int kq;   // COMPRESS row: nodes of its segment (1 or kCq); COPY row: 0; padding beyond n: -1
};
struct FusedSeg {
    int pb, pe;   // piece range of the segment
    int kb, kq;   // first node, node count
};
struct FusedPlan {
    std::vector<FusedRow> rows;    // [n rounded up to tile_eig]
    std::vector<FusedSeg> segs;    // the COMPRESS segments, eigen order
    std::vector<Segment> csegs;
    int npieces = 0, cnodes = 0;   // cnodes: nodes of COMPRESS segments (columns of G = U V per linear column)
};

inline void build_fused_plan(const CompressPlan& P, int tile_eig, int piece_eig, FusedPlan* F)
{
    const int padded = (P.n + tile_eig - 1) / tile_eig * tile_eig;
    F->rows.assign(padded, FusedRow{0, -1});
    F->segs.clear(); F->csegs.clear();
    F->npieces = 0; F->cnodes = 0;
    for (const Segment& sg : P.segs) {
        if (sg.type == kSegCompress) {
            FusedSeg sr{F->npieces, 0, sg.kb, sg.kq};
            int last = -1;
            for (int l = sg.l0; l < sg.l1; ++l) {
                if (l / piece_eig != last) { last = l / piece_eig; ++F->npieces; }
                F->rows[l] = FusedRow{F->npieces - 1, sg.kq};
            }
            sr.pe = F->npieces;
            F->segs.push_back(sr);
            F->csegs.push_back(sg);
            F->cnodes += sg.kq;
        } else {
            for (int l = sg.l0; l < sg.l1; ++l) F->rows[l] = FusedRow{sg.kb + (l - sg.l0), 0};
        }
    }
}

}  // namespace pg
