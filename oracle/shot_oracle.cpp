/*
 * shot_oracle.cpp -- CPU restatement of `shot.compute` / `shot.estimate_normal`.  TEST INFRASTRUCTURE ONLY.
 *
 * Reference call protocol: src_shot/shot.cpp:12-42 (estimate_normal) and :45-100 (compute): PCL
 * NormalEstimation (radius search, viewpoint at the origin) followed by SHOTEstimation<PointXYZ,
 * Normal, SHOT352> with its default SHOTLocalReferenceFrameEstimation, both at radius shot_r.
 *
 * Parity status: UNPINNED.  The arithmetic lives in PCL 1.9.1 (environment.yml:39, with flann 1.9.1
 * and Eigen), a third-party dependency that is neither vendored under the reference tree nor installed
 * in this image, and the reference ships no test, golden vector or stored descriptor for this path
 * (SURVEY.md section 8c).  What follows restates PCL 1.9.1's published algorithm function by function:
 *   features/normal_3d.h            NormalEstimation::computeFeature / computePointNormal,
 *                                   flipNormalTowardsViewpoint, solvePlaneParameters
 *   common/impl/centroid.hpp        computeMeanAndCovarianceMatrix (float, single pass, un-centred)
 *   common/impl/eigen.hpp           computeRoots, computeRoots2, eigen33 (smallest eigen-pair)
 *   features/impl/shot_lrf.hpp      SHOTLocalReferenceFrameEstimation::getLocalRF
 *   features/impl/shot.hpp          createBinDistanceShape, interpolateSingleChannel,
 *                                   normalizeHistogram, computePointSHOT, computeFeature
 *   kdtree/impl/kdtree_flann.hpp    radiusSearch: squared float distance strictly < float(r*r), self included
 * Deviations that cannot be avoided without PCL: the neighbour ORDER (kd-tree traversal order in PCL;
 * here ascending distance for the normals, ascending index for LRF/SHOT) which only perturbs float
 * accumulation order, and the symmetric 3x3 double eigen-solver (Eigen's tridiagonal QL there, cyclic
 * Jacobi here; both accurate to a few ulp of double).  It is self-checked against an independent
 * float64 numpy implementation in tests/test_shot_oracle.py.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {

const float kNaN = std::numeric_limits<float>::quiet_NaN();

struct Neighbor {
    int idx;
    float d2;
};

// Uniform grid over the cloud, cell edge = radius: a radius query touches the 27 surrounding cells.
struct CellGrid {
    float lo[3];
    float inv;
    int dim[3];
    std::vector<int> start;   // [cells+1]
    std::vector<int> order;   // point indices sorted by cell
    const float *pc;
    int64_t n;

    void build(const float *pts, int64_t count, double radius) {
        pc = pts;
        n = count;
        float hi[3];
        for (int k = 0; k < 3; ++k) lo[k] = hi[k] = pts[k];
        for (int64_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                float v = pts[3 * i + k];
                if (std::isfinite(v)) {
                    lo[k] = std::min(lo[k], v);
                    hi[k] = std::max(hi[k], v);
                }
            }
        inv = static_cast<float>(1.0 / radius);
        int64_t cells = 1;
        for (int k = 0; k < 3; ++k) {
            dim[k] = std::max(1, static_cast<int>((hi[k] - lo[k]) * inv) + 1);
            cells *= dim[k];
        }
        // pathological extents (radius tiny relative to the cloud): coarsen until the table is sane
        while (cells > (int64_t(1) << 26)) {
            inv *= 0.5f;
            cells = 1;
            for (int k = 0; k < 3; ++k) {
                dim[k] = std::max(1, static_cast<int>((hi[k] - lo[k]) * inv) + 1);
                cells *= dim[k];
            }
        }
        start.assign(cells + 1, 0);
        std::vector<int> cell_of(n);
        for (int64_t i = 0; i < n; ++i) {
            cell_of[i] = cell_index(pts + 3 * i);
            if (cell_of[i] >= 0) start[cell_of[i] + 1]++;
        }
        for (int64_t c = 0; c < cells; ++c) start[c + 1] += start[c];
        order.assign(start[cells], 0);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            if (cell_of[i] >= 0) order[fill[cell_of[i]]++] = static_cast<int>(i);
    }

    int coord(float v, int k) const {
        int c = static_cast<int>((v - lo[k]) * inv);
        return std::min(std::max(c, 0), dim[k] - 1);
    }

    int cell_index(const float *p) const {
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) return -1;
        return (coord(p[0], 0) * dim[1] + coord(p[1], 1)) * dim[2] + coord(p[2], 2);
    }

    // FLANN L2_Simple: float accumulation of diff*diff over x,y,z; RadiusResultSet keeps dist < radius^2
    void radius_search(const float *q, double radius, std::vector<Neighbor> &out) const {
        out.clear();
        if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) return;
        const float r2 = static_cast<float>(radius * radius);
        const int reach = std::max(1, static_cast<int>(std::ceil(radius * inv)));
        int c[3] = {coord(q[0], 0), coord(q[1], 1), coord(q[2], 2)};
        for (int x = std::max(c[0] - reach, 0); x <= std::min(c[0] + reach, dim[0] - 1); ++x)
            for (int y = std::max(c[1] - reach, 0); y <= std::min(c[1] + reach, dim[1] - 1); ++y)
                for (int z = std::max(c[2] - reach, 0); z <= std::min(c[2] + reach, dim[2] - 1); ++z) {
                    int cell = (x * dim[1] + y) * dim[2] + z;
                    for (int s = start[cell]; s < start[cell + 1]; ++s) {
                        const float *p = pc + 3 * order[s];
                        float d2 = 0.0f;
                        for (int k = 0; k < 3; ++k) {
                            float diff = q[k] - p[k];
                            d2 += diff * diff;
                        }
                        if (d2 < r2) out.push_back({order[s], d2});
                    }
                }
    }
};

// ---- pcl::computeRoots2 / computeRoots / eigen33 (float) ---------------------------------------------
void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = static_cast<float>(b * b - 4.0 * c);
    if (d < 0.0f) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

void compute_roots(const float m[9], float roots[3]) {
    // characteristic equation x^3 - c2 x^2 + c1 x - c0 = 0
    float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
    float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
    float c2 = m[0] + m[4] + m[8];
    if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = static_cast<float>(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

inline void cross3(const float a[3], const float b[3], float out[3]) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

// smallest eigenvalue / eigenvector of a symmetric 3x3 (row-major m), pcl::eigen33
void eigen33_smallest(const float mat[9], float &eigenvalue, float evec[3]) {
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat[i]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
    float s[9];
    for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
    float roots[3];
    compute_roots(s, roots);
    eigenvalue = roots[0] * scale;
    s[0] -= roots[0];
    s[4] -= roots[0];
    s[8] -= roots[0];
    float v1[3], v2[3], v3[3];
    cross3(s + 0, s + 3, v1);
    cross3(s + 0, s + 6, v2);
    cross3(s + 3, s + 6, v3);
    float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float *v;
    float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float inv = std::sqrt(l);
    for (int k = 0; k < 3; ++k) evec[k] = v[k] / inv;
}

// NormalEstimation::computePointNormal + flipNormalTowardsViewpoint(vp = origin)
void point_normal(const float *pc, const float *p, const std::vector<Neighbor> &nn, float normal[3]) {
    if (nn.size() < 3) {
        normal[0] = normal[1] = normal[2] = kNaN;
        return;
    }
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (const Neighbor &nb : nn) {
        const float *q = pc + 3 * nb.idx;
        accu[0] += q[0] * q[0];
        accu[1] += q[0] * q[1];
        accu[2] += q[0] * q[2];
        accu[3] += q[1] * q[1];
        accu[4] += q[1] * q[2];
        accu[5] += q[2] * q[2];
        accu[6] += q[0];
        accu[7] += q[1];
        accu[8] += q[2];
    }
    const float cnt = static_cast<float>(nn.size());
    for (int i = 0; i < 9; ++i) accu[i] /= cnt;
    float cov[9];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7];
    cov[5] = accu[4] - accu[7] * accu[8];
    cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
    float ev;
    eigen33_smallest(cov, ev, normal);
    // viewpoint (0,0,0): vp - point = -point
    float cos_theta = (-p[0]) * normal[0] + (-p[1]) * normal[1] + (-p[2]) * normal[2];
    if (cos_theta < 0) {
        normal[0] *= -1;
        normal[1] *= -1;
        normal[2] *= -1;
    }
}

// cyclic Jacobi for a symmetric 3x3 in double; eigenvalues ascending, eigenvectors in columns of V
void jacobi_eigen3(const double A_in[9], double w[3], double V[9]) {
    double A[9];
    std::memcpy(A, A_in, sizeof(A));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    double d[3] = {A[0], A[4], A[8]};
    std::sort(ord, ord + 3, [&](int a, int b) { return d[a] < d[b]; });
    double Vs[9];
    for (int j = 0; j < 3; ++j) {
        w[j] = d[ord[j]];
        for (int k = 0; k < 3; ++k) Vs[3 * k + j] = V[3 * k + ord[j]];
    }
    std::memcpy(V, Vs, sizeof(Vs));
}

// SHOTLocalReferenceFrameEstimation::getLocalRF; rf rows = x, y, z axes.  Returns false -> NaN frame.
bool local_rf(const float *pc, const float *p, const std::vector<Neighbor> &nn, double radius, float rf[9],
              int *margins = nullptr) {
    std::vector<double> vij;
    vij.reserve(nn.size() * 3);
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double sum = 0.0;
    int valid = 0;
    for (const Neighbor &nb : nn) {
        const float *q = pc + 3 * nb.idx;
        if (q[0] == p[0] && q[1] == p[1] && q[2] == p[2]) continue;
        double v[3] = {static_cast<double>(q[0] - p[0]), static_cast<double>(q[1] - p[1]), static_cast<double>(q[2] - p[2])};
        double distance = radius - std::sqrt(static_cast<double>(nb.d2));
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) cov[3 * a + b] += distance * (v[a] * v[b]);
        sum += distance;
        vij.insert(vij.end(), v, v + 3);
        ++valid;
    }
    if (valid < 5) {
        for (int i = 0; i < 9; ++i) rf[i] = kNaN;
        return false;
    }
    for (int i = 0; i < 9; ++i) cov[i] /= sum;
    double w[3], V[9];
    jacobi_eigen3(cov, w, V);
    if (!std::isfinite(w[0]) || !std::isfinite(w[1]) || !std::isfinite(w[2])) {
        for (int i = 0; i < 9; ++i) rf[i] = kNaN;
        return false;
    }
    double v1[3] = {V[2], V[5], V[8]};  // largest eigenvalue -> x
    double v3[3] = {V[0], V[3], V[6]};  // smallest -> z
    int plus_normal = 0, plus_tangent = 0;
    for (int ne = 0; ne < valid; ++ne) {
        const double *v = &vij[3 * ne];
        if (v[0] * v1[0] + v[1] * v1[1] + v[2] * v1[2] >= 0) ++plus_tangent;
        if (v[0] * v3[0] + v[1] * v3[1] + v[2] * v3[2] >= 0) ++plus_normal;
    }
    if (margins) {
        margins[0] = 2 * plus_tangent - valid;
        margins[1] = 2 * plus_normal - valid;
    }
    auto disambiguate = [&](int plus, double axis[3]) {
        plus = 2 * plus - valid;
        if (plus == 0) {
            const int points = 5, median = valid / 2;
            for (int i = -points / 2; i <= points / 2; ++i) {
                const double *v = &vij[3 * (median - i)];
                if (v[0] * axis[0] + v[1] * axis[1] + v[2] * axis[2] > 0) ++plus;
            }
            if (plus < points / 2 + 1)
                for (int k = 0; k < 3; ++k) axis[k] *= -1;
        } else if (plus < 0) {
            for (int k = 0; k < 3; ++k) axis[k] *= -1;
        }
    };
    disambiguate(plus_tangent, v1);
    disambiguate(plus_normal, v3);
    for (int k = 0; k < 3; ++k) {
        rf[k] = static_cast<float>(v1[k]);
        rf[6 + k] = static_cast<float>(v3[k]);
    }
    cross3(rf + 6, rf + 0, rf + 3);  // y = z x x (float)
    return true;
}

const double PST_PI = 3.1415926535897932384626433832795;
const double PST_RAD_45 = 0.78539816339744830961566084581988;
const double PST_RAD_90 = 1.5707963267948966192313216916398;
const double PST_RAD_135 = 2.3561944901923449288469825374596;
const double PST_RAD_PI_7_8 = 2.7488935718910690836548129603691;

// computePointSHOT: createBinDistanceShape + interpolateSingleChannel + normalizeHistogram
void point_shot(const float *pc, const float *normals, const float *p, const float rf[9],
                const std::vector<Neighbor> &nn, double radius, float shot[352]) {
    const int nr_bins = 10, max_sectors = 32, desc_len = 352;
    if (nn.size() < 5) {
        for (int i = 0; i < desc_len; ++i) shot[i] = kNaN;
        return;
    }
    const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
    const float *fx = rf, *fy = rf + 3, *fz = rf + 6;
    for (int i = 0; i < desc_len; ++i) shot[i] = 0.0f;
    for (const Neighbor &nb : nn) {
        const float *nrm = normals + 3 * nb.idx;
        if (!std::isfinite(nrm[0]) || !std::isfinite(nrm[1]) || !std::isfinite(nrm[2])) continue;
        double cosine = static_cast<double>(nrm[0] * fz[0] + nrm[1] * fz[1] + nrm[2] * fz[2] + 0.0f);  // Vector4f dot
        if (cosine > 1.0) cosine = 1.0;
        if (cosine < -1.0) cosine = -1.0;
        double bin_distance = ((1.0 + cosine) * nr_bins) / 2;

        const float *q = pc + 3 * nb.idx;
        float delta[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]};
        double distance = std::sqrt(static_cast<double>(nb.d2));
        if (std::fabs(distance - 0.0) < 1e-15) continue;
        double x_ref = static_cast<double>(delta[0] * fx[0] + delta[1] * fx[1] + delta[2] * fx[2] + 0.0f);
        double y_ref = static_cast<double>(delta[0] * fy[0] + delta[1] * fy[1] + delta[2] * fy[2] + 0.0f);
        double z_ref = static_cast<double>(delta[0] * fz[0] + delta[1] * fz[1] + delta[2] * fz[2] + 0.0f);
        if (std::fabs(y_ref) < 1E-30) y_ref = 0;
        if (std::fabs(x_ref) < 1E-30) x_ref = 0;
        if (std::fabs(z_ref) < 1E-30) z_ref = 0;

        unsigned char bit4 = ((y_ref > 0) || ((y_ref == 0.0) && (x_ref < 0))) ? 1 : 0;
        unsigned char bit3 = static_cast<unsigned char>(((x_ref > 0) || ((x_ref == 0.0) && (y_ref > 0))) ? !bit4 : bit4);
        int desc_index = (bit4 << 3) + (bit3 << 2);
        desc_index = desc_index << 1;
        if ((x_ref * y_ref > 0) || (x_ref == 0.0))
            desc_index += (std::fabs(x_ref) >= std::fabs(y_ref)) ? 0 : 4;
        else
            desc_index += (std::fabs(x_ref) > std::fabs(y_ref)) ? 4 : 0;
        desc_index += z_ref > 0 ? 1 : 0;
        desc_index += (distance > radius1_2) ? 2 : 0;

        int step_index = static_cast<int>(std::floor(bin_distance + 0.5));
        int volume_index = desc_index * (nr_bins + 1);
        bin_distance -= step_index;
        double weight = (1 - std::fabs(bin_distance));
        if (bin_distance > 0)
            shot[volume_index + ((step_index + 1) % nr_bins)] += static_cast<float>(bin_distance);
        else
            shot[volume_index + ((step_index - 1 + nr_bins) % nr_bins)] += -static_cast<float>(bin_distance);

        if (distance > radius1_2) {  // external sphere
            double rd = (distance - radius3_4) / radius1_2;
            if (distance > radius3_4)
                weight += 1 - rd;
            else {
                weight += 1 + rd;
                shot[(desc_index - 2) * (nr_bins + 1) + step_index] -= static_cast<float>(rd);
            }
        } else {  // internal sphere
            double rd = (distance - radius1_4) / radius1_2;
            if (distance < radius1_4)
                weight += 1 + rd;
            else {
                weight += 1 - rd;
                shot[(desc_index + 2) * (nr_bins + 1) + step_index] += static_cast<float>(rd);
            }
        }

        double incl_cos = z_ref / distance;
        if (incl_cos < -1.0) incl_cos = -1.0;
        if (incl_cos > 1.0) incl_cos = 1.0;
        double incl = std::acos(incl_cos);
        if (incl > PST_RAD_90 || (std::fabs(incl - PST_RAD_90) < 1e-30 && z_ref <= 0)) {
            double id = (incl - PST_RAD_135) / PST_RAD_90;
            if (incl > PST_RAD_135)
                weight += 1 - id;
            else {
                weight += 1 + id;
                shot[(desc_index + 1) * (nr_bins + 1) + step_index] -= static_cast<float>(id);
            }
        } else {
            double id = (incl - PST_RAD_45) / PST_RAD_90;
            if (incl < PST_RAD_45)
                weight += 1 + id;
            else {
                weight += 1 - id;
                shot[(desc_index - 1) * (nr_bins + 1) + step_index] += static_cast<float>(id);
            }
        }

        if (y_ref != 0.0 || x_ref != 0.0) {
            double azimuth = std::atan2(y_ref, x_ref);
            int sel = desc_index >> 2;
            double ad = (azimuth - (-PST_RAD_PI_7_8 + PST_RAD_45 * sel)) / PST_RAD_45;
            ad = std::max(-0.5, std::min(ad, 0.5));
            if (ad > 0) {
                weight += 1 - ad;
                int interp = (desc_index + 4) % max_sectors;
                shot[interp * (nr_bins + 1) + step_index] += static_cast<float>(ad);
            } else {
                int interp = (desc_index - 4 + max_sectors) % max_sectors;
                weight += 1 + ad;
                shot[interp * (nr_bins + 1) + step_index] -= static_cast<float>(ad);
            }
        }
        shot[volume_index + step_index] += static_cast<float>(weight);
    }
    double acc = 0.0;
    for (int j = 0; j < desc_len; ++j) acc += shot[j] * shot[j];
    acc = std::sqrt(acc);
    for (int j = 0; j < desc_len; ++j) shot[j] /= static_cast<float>(acc);
}

}  // namespace

// desc may be NULL (normals only, shot.cpp:12-42).  threads <= 1 reproduces the reference's
// single-threaded PCL classes; threads > 1 is the OpenMP steel-man used for the CPU baseline.
EXPORT int oracle_shot_compute(const float *pc, int64_t n, double normal_r, double shot_r, float *desc, float *normals,
                               int threads) {
    if (n < 0 || !pc || !normals) return 1;
    if (n == 0) return 0;
#ifdef _OPENMP
    const int nt = threads > 1 ? threads : 1;
#else
    const int nt = 1;
    (void)threads;
#endif
    (void)PST_PI;
    CellGrid grid_n;
    grid_n.build(pc, n, normal_r);
#pragma omp parallel num_threads(nt)
    {
        std::vector<Neighbor> nn;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            grid_n.radius_search(pc + 3 * i, normal_r, nn);
            // search::KdTree(sorted = true): ascending distance (shot.cpp:70-71)
            std::sort(nn.begin(), nn.end(), [](const Neighbor &a, const Neighbor &b) {
                return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
            });
            if (nn.empty()) {
                normals[3 * i] = normals[3 * i + 1] = normals[3 * i + 2] = kNaN;
            } else {
                point_normal(pc, pc + 3 * i, nn, normals + 3 * i);
            }
        }
    }
    if (!desc) return 0;
    CellGrid grid_s;
    grid_s.build(pc, n, shot_r);
#pragma omp parallel num_threads(nt)
    {
        std::vector<Neighbor> nn;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            const float *p = pc + 3 * i;
            float *out = desc + 352 * i;
            grid_s.radius_search(p, shot_r, nn);
            std::sort(nn.begin(), nn.end(), [](const Neighbor &a, const Neighbor &b) { return a.idx < b.idx; });
            float rf[9];
            bool ok = std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && local_rf(pc, p, nn, shot_r, rf) &&
                      !nn.empty();
            if (!ok) {
                for (int d = 0; d < 352; ++d) out[d] = kNaN;
                continue;
            }
            point_shot(pc, normals, p, rf, nn, shot_r, out);
        }
    }
    return 0;
}

// Exposes the local reference frames for tests ([n,9] rows x,y,z; NaN when invalid) and the sign-vote
// margins 2*plus - valid of the x and z axes ([n,2]; 0 = tie, resolved by search order in PCL).
EXPORT int oracle_shot_lrf(const float *pc, int64_t n, double shot_r, float *rf_out, int *margins_out) {
    CellGrid grid;
    grid.build(pc, n, shot_r);
    std::vector<Neighbor> nn;
    for (int64_t i = 0; i < n; ++i) {
        grid.radius_search(pc + 3 * i, shot_r, nn);
        std::sort(nn.begin(), nn.end(), [](const Neighbor &a, const Neighbor &b) { return a.idx < b.idx; });
        int m[2] = {0, 0};
        local_rf(pc, pc + 3 * i, nn, shot_r, rf_out + 9 * i, m);
        if (margins_out) {
            margins_out[2 * i] = m[0];
            margins_out[2 * i + 1] = m[1];
        }
    }
    return 0;
}
