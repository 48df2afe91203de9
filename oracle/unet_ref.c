/* CPU ORACLE, plain C (test infrastructure; never linked into the product).
 *
 * Independent restatement of the arithmetic behind
 * anatomix/model/network.py:530-548 (Unet.forward, standard branch) for 3-D
 * volumes.  The reference delegates every op to torch.nn layers
 * (torch==2.13.0, requirements.txt:10); the loops below restate the published
 * definitions of those layers:
 *   conv   : nn.Conv3d(k=3,s=1,padding='same',padding_mode='reflect')  network.py:310-318
 *   bn     : nn.BatchNorm3d in eval mode                                network.py:154-155
 *   inorm  : nn.InstanceNorm3d(affine=False), biased variance           network.py:157-158
 *   act    : ReLU / LeakyReLU(0.3)                                      network.py:188-191
 *   pool   : Max/AvgPool3d(2)                                           network.py:297,368
 *   up     : nn.Upsample(scale_factor=2, nearest|trilinear, align_corners=False) network.py:407
 *   concat : torch.cat((encoder, upsampled), dim=1)                     network.py:545
 * Layout: fp32 NCDHW, double accumulation inside the conv and the norms.
 * Pinned against tests/golden (outputs of the unmodified reference) by
 * tests/test_oracle.py.  Build: oracle/build_oracle.py (gcc -O2 -fopenmp).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

void ref_conv3d_reflect(const float *x, const float *w, const float *b, float *y,
                        int N, int Ci, int Co, int D, int H, int W)
{
    const long S = (long)D * H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int co = 0; co < Co; ++co)
            for (int z = 0; z < D; ++z)
                for (int yy = 0; yy < H; ++yy)
                    for (int xx = 0; xx < W; ++xx) {
                        double acc = b ? b[co] : 0.0;
                        for (int ci = 0; ci < Ci; ++ci) {
                            const float *xp = x + ((long)n * Ci + ci) * S;
                            const float *wp = w + ((long)co * Ci + ci) * 27;
                            for (int dz = 0; dz < 3; ++dz) {
                                int zz = reflect(z + dz - 1, D);
                                for (int dy = 0; dy < 3; ++dy) {
                                    int yv = reflect(yy + dy - 1, H);
                                    const float *row = xp + ((long)zz * H + yv) * W;
                                    const float *wr = wp + (dz * 3 + dy) * 3;
                                    acc += (double)wr[0] * row[reflect(xx - 1, W)]
                                         + (double)wr[1] * row[xx]
                                         + (double)wr[2] * row[reflect(xx + 1, W)];
                                }
                            }
                        }
                        y[((long)n * Co + co) * S + ((long)z * H + yy) * W + xx] = (float)acc;
                    }
}

void ref_batchnorm_eval(float *x, const float *gamma, const float *beta, const float *mean,
                        const float *var, float eps, int N, int C, long S)
{
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            float inv = 1.0f / sqrtf(var[c] + eps);
            float *p = x + ((long)n * C + c) * S;
            for (long i = 0; i < S; ++i) p[i] = (p[i] - mean[c]) * inv * gamma[c] + beta[c];
        }
}

void ref_instancenorm(float *x, float eps, int N, int C, long S)
{
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            float *p = x + ((long)n * C + c) * S;
            double s = 0, q = 0;
            for (long i = 0; i < S; ++i) s += p[i];
            double m = s / S;
            for (long i = 0; i < S; ++i) { double d = p[i] - m; q += d * d; }
            double inv = 1.0 / sqrt(q / S + eps);
            for (long i = 0; i < S; ++i) p[i] = (float)((p[i] - m) * inv);
        }
}

void ref_leaky_relu(float *x, float slope, long n)
{
    for (long i = 0; i < n; ++i) x[i] = x[i] > 0 ? x[i] : slope * x[i];
}

/* kind 0 = max, 1 = average */
void ref_pool2(const float *x, float *y, int kind, int NC, int D, int H, int W)
{
    int d = D / 2, h = H / 2, w = W / 2;
#pragma omp parallel for
    for (int c = 0; c < NC; ++c)
        for (int z = 0; z < d; ++z)
            for (int yy = 0; yy < h; ++yy)
                for (int xx = 0; xx < w; ++xx) {
                    float m = -INFINITY, s = 0;
                    for (int a = 0; a < 2; ++a)
                        for (int bq = 0; bq < 2; ++bq)
                            for (int cq = 0; cq < 2; ++cq) {
                                float v = x[(((long)c * D + 2 * z + a) * H + 2 * yy + bq) * W + 2 * xx + cq];
                                m = v > m ? v : m;
                                s += v;
                            }
                    y[(((long)c * d + z) * h + yy) * w + xx] = kind == 0 ? m : s * 0.125f;
                }
}

static inline void tri_src(int o, int n, int *i0, int *i1, float *t)
{   /* align_corners=False: src = (o+0.5)/2-0.5, clamped at 0 */
    float s = (o + 0.5f) * 0.5f - 0.5f;
    if (s < 0) s = 0;
    int a = (int)s;
    *i0 = a; *i1 = a + 1 < n ? a + 1 : n - 1; *t = s - a;
}

/* kind 0 = nearest, 1 = trilinear */
void ref_upsample2(const float *x, float *y, int kind, int NC, int D, int H, int W)
{
    int d = 2 * D, h = 2 * H, w = 2 * W;
#pragma omp parallel for
    for (int c = 0; c < NC; ++c) {
        const float *p = x + (long)c * D * H * W;
        for (int z = 0; z < d; ++z)
            for (int yy = 0; yy < h; ++yy)
                for (int xx = 0; xx < w; ++xx) {
                    float v;
                    if (kind == 0) v = p[((long)(z / 2) * H + yy / 2) * W + xx / 2];
                    else {
                        int z0, z1, y0, y1, x0, x1; float tz, ty, tx;
                        tri_src(z, D, &z0, &z1, &tz); tri_src(yy, H, &y0, &y1, &ty); tri_src(xx, W, &x0, &x1, &tx);
#define AT(a, b, c2) p[((long)(a) * H + (b)) * W + (c2)]
                        float c00 = AT(z0, y0, x0) * (1 - tx) + AT(z0, y0, x1) * tx;
                        float c01 = AT(z0, y1, x0) * (1 - tx) + AT(z0, y1, x1) * tx;
                        float c10 = AT(z1, y0, x0) * (1 - tx) + AT(z1, y0, x1) * tx;
                        float c11 = AT(z1, y1, x0) * (1 - tx) + AT(z1, y1, x1) * tx;
#undef AT
                        v = (c00 * (1 - ty) + c01 * ty) * (1 - tz) + (c10 * (1 - ty) + c11 * ty) * tz;
                    }
                    y[(((long)c * d + z) * h + yy) * w + xx] = v;
                }
    }
}

/* cat((a, b), dim=1) : a [N,Ca,S], b [N,Cb,S] -> y [N,Ca+Cb,S] */
void ref_concat(const float *a, const float *b, float *y, int N, int Ca, int Cb, long S)
{
    for (int n = 0; n < N; ++n) {
        memcpy(y + (long)n * (Ca + Cb) * S, a + (long)n * Ca * S, sizeof(float) * Ca * S);
        memcpy(y + ((long)n * (Ca + Cb) + Ca) * S, b + (long)n * Cb * S, sizeof(float) * Cb * S);
    }
}

/* Whole standard forward for the released topology (doubleconv, skip connections,
 * norm in {batch-eval(0), instance(1)}, act relu, pool in {max(0), avg(1)},
 * up in {nearest(0), trilinear(1)}).  Parameters arrive per conv in network order:
 * w[k] (Co,Ci,3,3,3), bias[k] or NULL, and for batch norm bn[k] = gamma|beta|mean|var
 * (4*Co floats) or NULL for the last conv.  Returns 0, or -1 on a bad shape. */
int ref_unet_forward(const float *x, float *out, int N, int D, int H, int W,
                     int input_nc, int output_nc, int num_downs, int ngf,
                     int norm_kind, float eps, int pool_kind, int up_kind,
                     const float *const *w, const float *const *bias, const float *const *bn)
{
    int unit = 1 << num_downs;
    if (D % unit || H % unit || W % unit || D < 2 * unit || H < 2 * unit || W < 2 * unit) return -1;
    float *skip[16]; int skipc[16];
    int k = 0, C = input_nc, d = D, h = H, wd = W;
    float *cur = (float *)malloc(sizeof(float) * N * C * (long)d * h * wd);
    memcpy(cur, x, sizeof(float) * N * C * (long)d * h * wd);

#define CONV_BLOCK(Cout, last) do {                                                        \
        long S_ = (long)d * h * wd;                                                        \
        float *nx_ = (float *)malloc(sizeof(float) * N * (Cout) * S_);                     \
        ref_conv3d_reflect(cur, w[k], bias ? bias[k] : NULL, nx_, N, C, (Cout), d, h, wd); \
        if (!(last)) {                                                                     \
            if (norm_kind == 0) ref_batchnorm_eval(nx_, bn[k], bn[k] + (Cout), bn[k] + 2 * (Cout), \
                                                   bn[k] + 3 * (Cout), eps, N, (Cout), S_);\
            else ref_instancenorm(nx_, eps, N, (Cout), S_);                                \
            ref_leaky_relu(nx_, 0.0f, (long)N * (Cout) * S_);                              \
        }                                                                                  \
        free(cur); cur = nx_; C = (Cout); ++k; } while (0)

    CONV_BLOCK(ngf, 0);
    for (int i = 0; i < num_downs; ++i) {
        int co = i == 0 ? C : 2 * C;
        CONV_BLOCK(co, 0);
        CONV_BLOCK(co, 0);
        skip[i] = cur; skipc[i] = C;
        float *p = (float *)malloc(sizeof(float) * N * C * (long)(d / 2) * (h / 2) * (wd / 2));
        ref_pool2(cur, p, pool_kind, N * C, d, h, wd);
        cur = p; d /= 2; h /= 2; wd /= 2;
    }
    { int co = 2 * C; CONV_BLOCK(co, 0); CONV_BLOCK(co, 0); }
    for (int j = num_downs - 1; j >= 0; --j) {
        float *u = (float *)malloc(sizeof(float) * N * C * (long)d * h * wd * 8);
        ref_upsample2(cur, u, up_kind, N * C, d, h, wd);
        free(cur); d *= 2; h *= 2; wd *= 2;
        long S = (long)d * h * wd;
        float *cat = (float *)malloc(sizeof(float) * N * (skipc[j] + C) * S);
        ref_concat(skip[j], u, cat, N, skipc[j], C, S);
        free(u); free(skip[j]);
        int half = C / 2;
        cur = cat; C = skipc[j] + C;
        CONV_BLOCK(half, 0);
        CONV_BLOCK(half, 0);
    }
    CONV_BLOCK(output_nc, 1);
    memcpy(out, cur, sizeof(float) * N * C * (long)d * h * wd);
    free(cur);
    return 0;
}
