// kgfit.cpp — hom/het k-mer classifier of the host glue: the grid-search branch of KGFitterBO::fit
// (reference src/Fitter.hpp:207-408, options as set at src/solver.cpp:777-785) and
// KmerGenieDiploidLike::classify (src/Classifier.hpp:59-80).
//
// All arithmetic is double and gates integer results (which colours are homozygous), so every value is
// produced by the same expression, in the same order, as in the reference; this file is
// compiled with the reference's flags.  What differs is only the schedule: the component
// densities do not depend on all eight parameters, so they are tabulated once per parameter subset
// (f_hom: 343 rows, f_het: 245 rows, f_err: 5 rows) instead of being recomputed for each of the 2.1 M grid
// points, and grid points are scanned in parallel with an ordered arg-min (smallest loop index among equal
// minima), which is exactly what the serial `if (nll < bestNLL)` keeps.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <limits>

#include "dgh.h"

namespace dgh {

namespace {

inline double derr_old_val(int c, double s) {           // Fitter.hpp:74-80, Classifier.hpp:111-118
    if (c <= 0) return 0.0;
    double a = std::pow((double)c, -s);
    double b = std::pow((double)(c + 1), -s);
    double v = a - b;
    return (v > 0.0 ? v : 1e-300);
}
inline std::vector<double> zeta_weights(double zp, int C) {   // Fitter.hpp:81-86
    std::vector<double> w(C + 1, 0.0);
    double S = 0.0;
    for (int k = 1; k <= C; k++) { w[k] = 1.0 / std::pow((double)k, zp); S += w[k]; }
    for (int k = 1; k <= C; k++) w[k] /= S;
    return w;
}
inline double normal_pdf(double x, double mu, double sd) {    // Fitter.hpp:87-91
    double s = std::max(sd, 1e-12), z = (x - mu) / s;
    static const double INV = 0.3989422804014327;
    return INV / s * std::exp(-0.5 * z * z);
}
inline double f_hom_x(int x, double u_v, double sd_v, const std::vector<double>& zeta, int C) {   // Fitter.hpp:103-111
    double sum = 0.0;
    for (int copy = 1; copy <= C; ++copy) {
        double mu = copy * u_v;
        double sd = std::sqrt((double)copy) * sd_v;
        sum += zeta[copy] * normal_pdf(x, mu, sd);
    }
    return std::max(sum, 1e-300);
}
inline double f_het_x(int x, double u_v, double var_w, const std::vector<double>& zeta, int C) {  // Fitter.hpp:112-122
    double u_base = 0.5 * u_v;
    double sd_base = 0.5 * std::sqrt(std::max(var_w, 1e-12));
    double sum = 0.0;
    for (int copy = 1; copy <= C; ++copy) {
        double mu = copy * u_base;
        double sd = std::sqrt((double)copy) * sd_base;
        sum += zeta[copy] * normal_pdf(x, mu, sd);
    }
    return std::max(sum, 1e-300);
}
inline std::vector<double> grid_or_freeze(double lo, double hi, int k) {   // Fitter.hpp:363-382
    if (std::fabs(hi - lo) < 1e-12) return {lo};
    std::vector<double> v;
    v.reserve(std::max(k, 1));
    if (k <= 1) { v.push_back((lo + hi) / 2.0); return v; }
    for (int i = 0; i < k; i++) {
        double t = (double)i / (double)(k - 1);
        v.push_back(lo + t * (hi - lo));
    }
    return v;
}

}  // namespace

KGFit kg_fit(const std::vector<std::pair<int, double>>& hist, int max_multiplicity, int threads) {
    // options: src/solver.cpp:777-785 on top of the defaults of src/Fitter.hpp:25-46
    const int max_copy = 10, max_x_use = max_multiplicity;
    const double u_lo = 1, u_hi = max_multiplicity, sd_lo = 0.5, sd_hi = 2.0, varw_lo = 0.71, varw_hi = 4.0;
    const double pd_lo = 0.1, pd_hi = 1.0, pe_lo = 0.0, pe_hi = 0.1, s_lo = 1.01, s_hi = 4.0, zp_lo = 1.01, zp_hi = 4.0;
    const int grid_u = 7, grid_sd = 7, grid_varw = 5, grid_pd = 7, grid_pe = 5, grid_s = 5, grid_zp = 7;

    // dense histogram 0..N (Fitter.hpp:209-213)
    int Nmax = 0;
    for (auto& b : hist) Nmax = std::max(Nmax, b.first);
    const int N = std::min(Nmax, max_x_use);
    std::vector<double> H(N + 1, 0.0);
    for (auto& b : hist) if (b.first <= N && b.first >= 0) H[b.first] += b.second;

    const auto U = grid_or_freeze(u_lo, u_hi, grid_u), SD = grid_or_freeze(sd_lo, sd_hi, grid_sd);
    const auto VW = grid_or_freeze(varw_lo, varw_hi, grid_varw), ZP = grid_or_freeze(zp_lo, zp_hi, grid_zp);
    const auto ZPH = grid_or_freeze(zp_lo, zp_hi, grid_zp), PD = grid_or_freeze(pd_lo, pd_hi, grid_pd);
    const auto PE = grid_or_freeze(pe_lo, pe_hi, grid_pe), SS = grid_or_freeze(s_lo, s_hi, grid_s);
    const size_t nU = U.size(), nSD = SD.size(), nVW = VW.size(), nZP = ZP.size(), nZPH = ZPH.size(), nPD = PD.size(),
                 nPE = PE.size(), nS = SS.size();

    // bins that contribute (nll_hist skips y <= 0, Fitter.hpp:134-135), ascending x like the reference loop
    std::vector<int> xs;
    const int Nuse = std::min((int)H.size() - 1, max_x_use);
    for (int x = 1; x <= Nuse; ++x) if (H[x] > 0) xs.push_back(x);
    const size_t nx = xs.size();

    // component tables
    std::vector<std::vector<double>> zeta_zp(nZP), zeta_zph(nZPH);
    for (size_t a = 0; a < nZP; ++a) zeta_zp[a] = zeta_weights(ZP[a], max_copy);
    for (size_t a = 0; a < nZPH; ++a) zeta_zph[a] = zeta_weights(ZPH[a], max_copy);
    std::vector<double> FH(nU * nSD * nZP * nx), FT(nU * nVW * nZPH * nx), FE(nS * nx);
    for (size_t iu = 0; iu < nU; ++iu)
        for (size_t is = 0; is < nSD; ++is)
            for (size_t iz = 0; iz < nZP; ++iz)
                for (size_t b = 0; b < nx; ++b)
                    FH[((iu * nSD + is) * nZP + iz) * nx + b] = f_hom_x(xs[b], U[iu], SD[is], zeta_zp[iz], max_copy);
    for (size_t iu = 0; iu < nU; ++iu)
        for (size_t iv = 0; iv < nVW; ++iv)
            for (size_t iz = 0; iz < nZPH; ++iz)
                for (size_t b = 0; b < nx; ++b)
                    FT[((iu * nVW + iv) * nZPH + iz) * nx + b] = f_het_x(xs[b], U[iu], VW[iv], zeta_zph[iz], max_copy);
    for (size_t is = 0; is < nS; ++is)
        for (size_t b = 0; b < nx; ++b) FE[is * nx + b] = derr_old_val(xs[b], SS[is]);

    // grid scan, loop order u, sd, vw, zp, zph, pd, pe, s (Fitter.hpp:391-405)
    const size_t outer = nU * nSD * nVW;
    std::vector<double> best_nll(outer, std::numeric_limits<double>::infinity());
    std::vector<size_t> best_inner(outer, 0);
    std::vector<char> best_set(outer, 0);
    const size_t inner = nZP * nZPH * nPD * nPE * nS;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (long long o = 0; o < (long long)outer; ++o) {
        const size_t iu = (size_t)o / (nSD * nVW), is = ((size_t)o / nVW) % nSD, iv = (size_t)o % nVW;
        double bn = std::numeric_limits<double>::infinity();
        size_t bi = 0;
        bool set = false;
        size_t idx = 0;
        for (size_t izp = 0; izp < nZP; ++izp) {
            const double* fh = &FH[((iu * nSD + is) * nZP + izp) * nx];
            for (size_t izh = 0; izh < nZPH; ++izh) {
                const double* ft = &FT[((iu * nVW + iv) * nZPH + izh) * nx];
                for (size_t ipd = 0; ipd < nPD; ++ipd)
                    for (size_t ipe = 0; ipe < nPE; ++ipe)
                        for (size_t isx = 0; isx < nS; ++isx, ++idx) {
                            const double p_d = PD[ipd], p_e = PE[ipe];
                            const double* fe_t = &FE[isx * nx];
                            double nll = 0.0;
                            for (size_t b = 0; b < nx; ++b) {
                                double y = H[xs[b]];
                                double fe = fe_t[b];
                                double fhet = ft[b];
                                double fhom = fh[b];
                                double mix = p_e * fe + (1.0 - p_e) * (p_d * fhet + (1.0 - p_d) * fhom);   // Fitter.hpp:139-140
                                nll += -y * std::log(mix + 1e-300);                                        // :141
                            }
                            if (nll < bn) { bn = nll; bi = idx; set = true; }
                        }
            }
        }
        best_nll[o] = bn; best_inner[o] = bi; best_set[o] = set ? 1 : 0;
    }
    double bestNLL = std::numeric_limits<double>::infinity();
    size_t bo = 0, bi = 0;
    bool any = false;
    for (size_t o = 0; o < outer; ++o)
        if (best_set[o] && best_nll[o] < bestNLL) { bestNLL = best_nll[o]; bo = o; bi = best_inner[o]; any = true; }

    KGFit r;
    r.P.max_copy = max_copy;
    if (any) {
        const size_t iu = bo / (nSD * nVW), is = (bo / nVW) % nSD, iv = bo % nVW;
        size_t t = bi;
        const size_t isx = t % nS; t /= nS;
        const size_t ipe = t % nPE; t /= nPE;
        const size_t ipd = t % nPD; t /= nPD;
        const size_t izh = t % nZPH; t /= nZPH;
        const size_t izp = t;
        r.P.u_v = U[iu]; r.P.sd_v = SD[is]; r.P.var_w = VW[iv]; r.P.zp_copy = ZP[izp]; r.P.zp_copy_het = ZPH[izh];
        r.P.p_d = PD[ipd]; r.P.p_e = PE[ipe]; r.P.err_shape = SS[isx];
    } else {
        // no grid point improved on +inf (e.g. an empty histogram): the reference keeps its seed P0
        // (Fitter.hpp:236-246); seeds need the valley/peak heuristics, which index H[2] and are undefined
        // for such inputs, so the documented defaults are used instead.
        r.P.u_v = 4.0; r.P.sd_v = 1.2; r.P.var_w = 2.0; r.P.zp_copy = 1.5; r.P.zp_copy_het = 1.7; r.P.p_d = 0.05;
        r.P.p_e = 0.05; r.P.err_shape = 2.0;
    }
    r.nll = bestNLL;
    return r;
}

bool kg_is_hom(const KGParams& P, int x) {                       // Classifier.hpp:59-80
    const std::vector<double> zeta_hom = zeta_weights(P.zp_copy, P.max_copy);
    const std::vector<double> zeta_het = zeta_weights(P.zp_copy_het, P.max_copy);
    double fe = derr_old_val(x, P.err_shape);
    double fhet = f_het_x(x, P.u_v, P.var_w, zeta_het, P.max_copy);   // valHet_ (:157-168) is the same expression
    double fhom = f_hom_x(x, P.u_v, P.sd_v, zeta_hom, P.max_copy);    // valHom_ (:147-155)
    double a = P.p_e * fe;
    double b = (1.0 - P.p_e) * P.p_d * fhet;
    double c = (1.0 - P.p_e) * (1.0 - P.p_d) * fhom;
    double Z = std::max(a + b + c, 1e-300);
    double phet = b / Z, phom = c / Z;
    (void)a;
    return !(x == 1 || phet >= phom);
}

}  // namespace dgh
