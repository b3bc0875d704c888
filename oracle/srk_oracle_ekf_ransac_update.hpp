// TEST INFRASTRUCTURE ONLY — CPU oracle: the two-stage 1-point RANSAC update of suriko-engine's MonoSLAM
// ("EKF.cpp" = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp), the update the shipped flagfile selects
// (--monoslam_update_impl=4, flagfile-demo-davison-mono-slam.txt:33).
//   EkfProjectedCovariance        GetSalientPointProjected2DPosWithUncertainty            EKF.cpp:3901-4030 (J P_in J^T, FixAlmostSymmetricMat,
//                                 CheckEllipseIsExtractableFrom2DCovarMat :505-512 -> obs-geom.cpp:873-905: no negative eigenvalue beyond IsClose)
//   EkfOnePointRansacUpdate       ProcessFrame_OnePointRansacUpdateCore                   EKF.cpp:1393-1513
//        stage 1: consensus set of the best 1-point hypothesis (:1433-1434) -> stacked update with the low-innovation inliers (:1445-1446)
//        stage 2: every other matched point whose corner lies inside the chi^2 (dof 2, 99 %: the float literal 9.21034f) ellipse of its
//                 projection UNDER THE UPDATED STATE is rescued (:1466-1498) -> second stacked update (:1507-1508)
// Parity unpinned (no reference test for this path).
#pragma once
#include <cmath>
#include <vector>
#include "srk_oracle_ekf.hpp"
#include "srk_oracle_ekf_ransac.hpp"

namespace srk_oracle {

// covar2D = J P_in J^T, J = [d hd / d (camera position, quaternion) | d hd / d point], P_in = the matching blocks of P.  Returns false when
// the ellipse check fails.  Hx [2 x 13], Hy [2 x s] row-major.
inline bool EkfProjectedCovariance(const EkfMat& P, const double* Hx, const double* Hy, size_t off, int s, double cov[2][2]) {
    const int d = 7 + s;
    std::vector<double> J(2 * (size_t)d), Pin((size_t)d * d);
    auto idx = [&](int i) { return i < 7 ? (size_t)i : off + (size_t)(i - 7); };
    for (int k = 0; k < 2; ++k) { for (int c = 0; c < 7; ++c) J[(size_t)k * d + c] = Hx[k * 13 + c]; for (int c = 0; c < s; ++c) J[(size_t)k * d + 7 + c] = Hy[k * s + c]; }
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) Pin[(size_t)i * d + j] = P(idx(i), idx(j));
    double c2[2][2];
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            double acc = 0;
            for (int i = 0; i < d; ++i) { double t = 0; for (int j = 0; j < d; ++j) t += Pin[(size_t)i * d + j] * J[(size_t)b * d + j]; acc += J[(size_t)a * d + i] * t; }
            c2[a][b] = acc;
        }
    cov[0][0] = c2[0][0]; cov[1][1] = c2[1][1]; cov[0][1] = cov[1][0] = (c2[0][1] + c2[1][0]) / 2;    // FixAlmostSymmetricMat
    // eigenvalues of the symmetric 2x2, ascending; a negative one is tolerated only when IsClose(0, value) (obs-geom.cpp:895-905)
    const double tr = cov[0][0] + cov[1][1], df = cov[0][0] - cov[1][1];
    const double rad = std::sqrt(df * df / 4 + cov[0][1] * cov[0][1]);
    const double lo = tr / 2 - rad;
    if (lo < 0 && !IsClose<double>(0.0, lo)) return false;
    return true;
}

// Returns (low-innovation inliers, rescued high-innovation points) in counts[2]; low[m] / high[m] = the two masks.  x, P are updated in place.
inline void EkfOnePointRansacUpdate(std::vector<double>* x, EkfMat* P, int64_t m, const int64_t* pt_off, int s, const double* z, double meas_var, const EkfCamera& cam,
                                    double max_divergence_pix, double chi2_thr, unsigned char* low, unsigned char* high, int64_t counts[2]) {
    const size_t n = x->size();
    auto jacobians = [&](const std::vector<int64_t>& sel, std::vector<double>* Hc, std::vector<double>* Hp, std::vector<double>* hp) {
        Hc->assign(sel.size() * 2 * 13, 0.0); Hp->assign(sel.size() * 2 * (size_t)s, 0.0); hp->assign(sel.size() * 2, 0.0);
        for (size_t i = 0; i < sel.size(); ++i)
            EkfMeasurementJacobian(cam, x->data(), x->data() + pt_off[sel[i]], s, Hc->data() + i * 26, Hp->data() + i * 2 * (size_t)s, hp->data() + 2 * i);
    };
    auto stacked = [&](const std::vector<int64_t>& sel, const std::vector<double>& Hc, const std::vector<double>& Hp, const std::vector<double>& hp) {
        EkfMat H(2 * sel.size(), n);
        std::vector<double> zs(2 * sel.size());
        for (size_t i = 0; i < sel.size(); ++i)
            for (int k = 0; k < 2; ++k) {
                const size_t row = 2 * i + (size_t)k;
                for (int c = 0; c < 13; ++c) H(row, (size_t)c) = Hc[row * 13 + c];
                for (int c = 0; c < s; ++c) H(row, (size_t)(pt_off[sel[i]] + c)) = Hp[row * (size_t)s + c];
                zs[row] = z[2 * sel[i] + k];
            }
        EkfStackedUpdate(x, P, H, zs, hp, meas_var, true);
    };
    std::vector<int64_t> all((size_t)m);
    for (int64_t i = 0; i < m; ++i) all[(size_t)i] = i;
    std::vector<double> Hc, Hp, hp;
    jacobians(all, &Hc, &Hp, &hp);
    std::vector<int32_t> support((size_t)m);
    std::vector<unsigned char> inl((size_t)m, 0);
    const int best = EkfRansacConsensus(*x, *P, m, Hc.data(), Hp.data(), pt_off, s, z, meas_var, cam, max_divergence_pix, support.data(), inl.data());
    std::vector<int64_t> lows;
    for (int64_t i = 0; i < m; ++i) { low[i] = (best >= 0 && inl[(size_t)i]) ? 1 : 0; high[i] = 0; if (low[i]) lows.push_back(i); }
    counts[0] = (int64_t)lows.size(); counts[1] = 0;
    if (!lows.empty()) {
        std::vector<double> Hc1, Hp1, hp1;
        jacobians(lows, &Hc1, &Hp1, &hp1);         // the stacked update derives at the state it is handed (the predicted one: :1446, :987-1003)
        stacked(lows, Hc1, Hp1, hp1);
    }
    if ((int64_t)lows.size() == m) return;          // no candidates to rescue (:1468-1470)
    std::vector<int64_t> rescued;
    for (int64_t i = 0; i < m; ++i) {
        if (low[i]) continue;
        double Hx[26], Hy[12], hd[2];
        EkfMeasurementJacobian(cam, x->data(), x->data() + pt_off[i], s, Hx, Hy, hd);
        double cov[2][2];
        if (!EkfProjectedCovariance(*P, Hx, Hy, (size_t)pt_off[i], s, cov)) continue;   // the reference asserts op_cov; a failed ellipse is not rescued here
        const double det = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
        const double i00 = cov[1][1] / det, i01 = -cov[0][1] / det, i10 = -cov[1][0] / det, i11 = cov[0][0] / det;    // Eigen's 2x2 inverse: cofactors / determinant
        const double d0 = z[2 * i] - hd[0], d1 = z[2 * i + 1] - hd[1];
        const double dist = d0 * (i00 * d0 + i01 * d1) + d1 * (i10 * d0 + i11 * d1);
        if (dist < chi2_thr) { high[i] = 1; rescued.push_back(i); }
    }
    counts[1] = (int64_t)rescued.size();
    if (!rescued.empty()) {
        std::vector<double> Hc2, Hp2, hp2;
        jacobians(rescued, &Hc2, &Hp2, &hp2);
        stacked(rescued, Hc2, Hp2, hp2);
    }
}

}  // namespace srk_oracle
