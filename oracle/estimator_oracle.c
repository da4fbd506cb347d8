/* ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C FP64 restatement of the reference's transformation estimators:
 *   E2  Threepps2Tran  tests/estimator.py:55-63  (Kabsch from 3 matches, NO reflection fix)
 *   E3  overlap_cal    tests/estimator.py:66-70  + transform_points utils/utils.py:42-50
 *   E4  yohoc.ransac   tests/estimator.py:119-137 (loop over pre-drawn hypotheses, strict '>')
 *   E5  yohoo.ransac   tests/estimator.py:330-336
 *
 * np.linalg.svd is LAPACK dgesdd (third-party, not under /root/reference; numpy is un-pinned in the
 * reference's requirements.txt).  With three matches the 3x3 cross-covariance H has rank <= 2, so the
 * third singular pair is a null-space pair whose SIGN is implementation noise in LAPACK (measured here:
 * det(VT.T@U.T) = +1 in 48.9 % of 2000 well-posed triplets).  The restatement therefore fixes an explicit,
 * reproducible algorithm (DESIGN.md "Estimator arithmetic"):
 *     one-sided (Hestenes) Jacobi SVD, 8 cyclic sweeps, pairs (0,1),(0,2),(1,2);
 *     u3 = u1 x u2, v3 = v1 x v2;  R = v1 u1^T + v2 u2^T + s * v3 u3^T,
 *     s = sign of det(H) evaluated in FP64 by cofactor expansion along row 0 (s=+1 when det==0),
 *     or the caller's override (+1/-1) — used to replay the sign LAPACK happened to produce.
 * All arithmetic is IEEE FP64 with the operation order written below; compile with -ffp-contract=off.
 * fma() is used only where written.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o oracle/_build/libestimator_oracle.so oracle/estimator_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static double dot3(const double a[3], const double b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/* Column pair rotation of the one-sided Jacobi iteration on W (3x3, column j = W[:,j]) and V. */
static void hestenes_pair(double W[3][3], double V[3][3], int p, int q) {
    double wp[3] = {W[0][p], W[1][p], W[2][p]};
    double wq[3] = {W[0][q], W[1][q], W[2][q]};
    double alpha = dot3(wp, wp), beta = dot3(wq, wq), gamma = dot3(wp, wq);
    if (gamma == 0.0) return;
    double zeta = (beta - alpha) / (2.0 * gamma);
    double az = fabs(zeta);
    double t = 1.0 / (az + sqrt(1.0 + zeta * zeta));
    if (zeta < 0.0) t = -t;
    double c = 1.0 / sqrt(1.0 + t * t);
    double s = c * t;
    for (int r = 0; r < 3; ++r) {
        double a = W[r][p], b = W[r][q];
        W[r][p] = c * a - s * b;
        W[r][q] = s * a + c * b;
        double va = V[r][p], vb = V[r][q];
        V[r][p] = c * va - s * vb;
        V[r][q] = s * va + c * vb;
    }
}

/* Kabsch from three matches; returns 1 when the triplet is degenerate (rank(H) < 2). */
int yoho_oracle_kabsch3(const double* k0, const double* k1, const int32_t ids[3], int sign_override,
                        double T[12]) {
    double c0[3], c1[3], a[3][3], b[3][3], H[3][3];
    for (int d = 0; d < 3; ++d) {
        c0[d] = ((k0[3 * ids[0] + d] + k0[3 * ids[1] + d]) + k0[3 * ids[2] + d]) / 3.0;
        c1[d] = ((k1[3 * ids[0] + d] + k1[3 * ids[1] + d]) + k1[3 * ids[2] + d]) / 3.0;
    }
    for (int i = 0; i < 3; ++i)
        for (int d = 0; d < 3; ++d) {
            a[i][d] = k1[3 * ids[i] + d] - c1[d];   /* kps1 - center1 */
            b[i][d] = k0[3 * ids[i] + d] - c0[d];   /* kps0 - center0 */
        }
    /* H = (kps1-c1)^T (kps0-c0): H[r][c] = sum_i a[i][r] b[i][c] */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            H[r][c] = fma(a[2][r], b[2][c], fma(a[1][r], b[1][c], a[0][r] * b[0][c]));

    double s;
    if (sign_override > 0) s = 1.0;
    else if (sign_override < 0) s = -1.0;
    else {
        double m0 = H[1][1] * H[2][2] - H[1][2] * H[2][1];
        double m1 = H[1][0] * H[2][2] - H[1][2] * H[2][0];
        double m2 = H[1][0] * H[2][1] - H[1][1] * H[2][0];
        double det = (H[0][0] * m0 - H[0][1] * m1) + H[0][2] * m2;
        s = det < 0.0 ? -1.0 : 1.0;
    }

    double W[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    memcpy(W, H, sizeof(W));
    for (int sweep = 0; sweep < 8; ++sweep) {
        hestenes_pair(W, V, 0, 1);
        hestenes_pair(W, V, 0, 2);
        hestenes_pair(W, V, 1, 2);
    }
    double n2[3];
    for (int j = 0; j < 3; ++j) {
        double w[3] = {W[0][j], W[1][j], W[2][j]};
        n2[j] = dot3(w, w);
    }
    /* two largest columns, stable (lowest index wins ties) */
    int i1 = 0;
    if (n2[1] > n2[i1]) i1 = 1;
    if (n2[2] > n2[i1]) i1 = 2;
    int i2 = -1;
    for (int j = 0; j < 3; ++j) {
        if (j == i1) continue;
        if (i2 < 0 || n2[j] > n2[i2]) i2 = j;
    }
    double R[3][3];
    int degenerate = 0;
    if (n2[i1] == 0.0) {
        degenerate = 1;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = (r == c) ? 1.0 : 0.0;
    } else {
        double s1 = sqrt(n2[i1]);
        double u1[3] = {W[0][i1] / s1, W[1][i1] / s1, W[2][i1] / s1};
        double v1[3] = {V[0][i1], V[1][i1], V[2][i1]};
        double v2[3] = {V[0][i2], V[1][i2], V[2][i2]};
        double u2[3];
        if (n2[i2] <= 1e-28 * n2[i1]) {
            degenerate = 1;
            int j = 0;
            if (fabs(u1[1]) < fabs(u1[j])) j = 1;
            if (fabs(u1[2]) < fabs(u1[j])) j = 2;
            for (int d = 0; d < 3; ++d) u2[d] = ((d == j) ? 1.0 : 0.0) - u1[j] * u1[d];
        } else {
            double s2 = sqrt(n2[i2]);
            for (int d = 0; d < 3; ++d) u2[d] = W[d][i2] / s2;
            double pr = dot3(u1, u2);
            for (int d = 0; d < 3; ++d) u2[d] = u2[d] - pr * u1[d];
        }
        double nn = sqrt(dot3(u2, u2));
        for (int d = 0; d < 3; ++d) u2[d] = u2[d] / nn;
        double u3[3], v3[3];
        cross3(u1, u2, u3);
        cross3(v1, v2, v3);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                R[r][c] = (v1[r] * u1[c] + v2[r] * u2[c]) + s * (v3[r] * u3[c]);
    }
    for (int r = 0; r < 3; ++r) {
        T[4 * r + 0] = R[r][0]; T[4 * r + 1] = R[r][1]; T[4 * r + 2] = R[r][2];
        /* offset = center0 - center1 @ R^T */
        T[4 * r + 3] = c0[r] - fma(R[r][2], c1[2], fma(R[r][1], c1[1], R[r][0] * c1[0]));
    }
    return degenerate;
}

/* E3: number of matches with || k0 - (R k1 + t) ||^2 < thr2 (strict).  mask may be NULL. */
int32_t yoho_oracle_count_inliers(const double* k0, const double* k1, int32_t M, const double T[12],
                                  double thr2, uint8_t* mask) {
    int32_t n = 0;
    for (int32_t m = 0; m < M; ++m) {
        double x = k1[3 * m], y = k1[3 * m + 1], z = k1[3 * m + 2];
        double px = fma(T[2], z, fma(T[1], y, T[0] * x)) + T[3];
        double py = fma(T[6], z, fma(T[5], y, T[4] * x)) + T[7];
        double pz = fma(T[10], z, fma(T[9], y, T[8] * x)) + T[11];
        double dx = k0[3 * m] - px, dy = k0[3 * m + 1] - py, dz = k0[3 * m + 2] - pz;
        double diff = fma(dz, dz, fma(dy, dy, dx * dx));
        int in = diff < thr2;
        if (mask) mask[m] = (uint8_t)in;
        n += in;
    }
    return n;
}

/* E4 over a pre-drawn hypothesis list.  signs may be NULL (rule) or int8[iters] in {-1,0,+1,2}; 2 = take the caller's
 * transform fixed[it] (float64 [iters,12]) for that hypothesis instead of the Kabsch solve.
 * counts/degen may be NULL.  best_iter = -1 and T = [I|0] when no hypothesis has an inlier. */
void yoho_oracle_yohoc(const double* k0, const double* k1, int32_t M, const int32_t* hyp, int32_t iters,
                       const int8_t* signs, const double* fixed, double dist, double T_best[12], int32_t* best_iter,
                       int32_t* n_inl, uint8_t* mask, int32_t* counts, uint8_t* degen) {
    double thr2 = dist * dist;
    int32_t best = 0, bi = -1;
    double T[12];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T_best[4 * r + c] = (r == c) ? 1.0 : 0.0;
    for (int32_t it = 0; it < iters; ++it) {
        int dg = yoho_oracle_kabsch3(k0, k1, hyp + 3 * it, signs ? signs[it] : 0, T);
        if (signs && fixed && signs[it] == 2) memcpy(T, fixed + 12 * (size_t)it, sizeof(T));
        int32_t n = yoho_oracle_count_inliers(k0, k1, M, T, thr2, 0);
        if (counts) counts[it] = n;
        if (degen) degen[it] = (uint8_t)dg;
        if (n > best) { best = n; bi = it; memcpy(T_best, T, sizeof(T)); }
    }
    *best_iter = bi;
    *n_inl = best;
    if (mask) {
        if (bi >= 0) yoho_oracle_count_inliers(k0, k1, M, T_best, thr2, mask);
        else memset(mask, 0, (size_t)M);
    }
}

/* E5: score given 3x4 hypotheses in order, keep the first strictly-best. */
void yoho_oracle_yohoo(const double* k0, const double* k1, int32_t M, const double* trans, int32_t H,
                       double dist, double T_best[12], int32_t* best_iter, int32_t* n_inl, uint8_t* mask,
                       int32_t* counts) {
    double thr2 = dist * dist;
    int32_t best = 0, bi = -1;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T_best[4 * r + c] = (r == c) ? 1.0 : 0.0;
    for (int32_t h = 0; h < H; ++h) {
        int32_t n = yoho_oracle_count_inliers(k0, k1, M, trans + 12 * h, thr2, 0);
        if (counts) counts[h] = n;
        if (n > best) { best = n; bi = h; memcpy(T_best, trans + 12 * h, 12 * sizeof(double)); }
    }
    *best_iter = bi;
    *n_inl = best;
    if (mask) {
        if (bi >= 0) yoho_oracle_count_inliers(k0, k1, M, T_best, thr2, mask);
        else memset(mask, 0, (size_t)M);
    }
}
