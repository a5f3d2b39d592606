/* CPU restatement (plain C, float64) of the reference hot loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
 * path and the timed CPU baseline of bench.py.  Never linked into the product.
 *
 * orc_wc_run   follows /root/reference/netwWilsonCowanPlastic.py:77-83 (wilsonCowan)
 *              and :86-137 (run): explicit Euler-Maruyama, three phases with
 *              tau_ip = {0.05, 1, 2}, state stored BEFORE the update whenever
 *              i % downsamp == 0 in phase 3.
 * orc_bold_sim follows the Balloon-Windkessel restatement of oracle/bold_oracle.py
 *              (BOLDModel.Sim, call site netwWilsonCowanPlastic.py:144; parity unpinned).
 * orc_philox_normals  the "nrem-philox-v2" stream of oracle/philox.py (Philox4x32-7).
 *
 * Build: make -C oracle   (gcc -O2 -fno-fast-math; scalar, one thread)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double a_ee, a_ie_0, a_ei, a_ii;
    double tauE, tauI;
    double P, rhoE, rE, rI, mu, sigmaI;
    double dtSim, sqdtD;
    double E0, I0;
    double tau_ip[3];
} orc_wc_params;

static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
}

#define ORC_PHILOX_ROUNDS 7       /* the stream definition (oracle/philox.py ROUNDS) */

void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

static inline double u23(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

/* z[0..N) standard normals of (seed, stream, step) */
void orc_philox_normals(uint64_t seed, uint64_t stream, uint32_t step, int N, double* z) {
    const double two_pi = 6.283185307179586476925286766559;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int nq = (N + 3) / 4;
    for (int q = 0; q < nq; ++q) {
        uint32_t ctr[4] = {step, (uint32_t)q, (uint32_t)stream, (uint32_t)(stream >> 32)};
        uint32_t x[4];
        double v[4];
        orc_philox4x32(ctr, key, ORC_PHILOX_ROUNDS, x);
        double r0 = sqrt(-2.0 * log(u23(x[0]))), a0 = two_pi * (u23(x[1]) - 0.5);
        double r1 = sqrt(-2.0 * log(u23(x[2]))), a1 = two_pi * (u23(x[3]) - 0.5);
        v[0] = r0 * cos(a0); v[1] = r0 * sin(a0);
        v[2] = r1 * cos(a1); v[3] = r1 * sin(a1);
        for (int j = 0; j < 4 && 4 * q + j < N; ++j) z[4 * q + j] = v[j];
    }
}

/* Fast CPU noise for the timed baseline only (bench.py): xoshiro256** + Marsaglia polar method, the same
 * algorithm class as the reference's generator (numba: MT19937 + polar Gaussian).  Never used for parity. */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(uint64_t s[4]) {
    const uint64_t r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static inline void polar_normals(uint64_t st[4], int N, double scale, double* z) {
    int i = 0;
    while (i < N) {
        double u, v, r2;
        do {
            u = 2.0 * ((double)(xo_next(st) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
            v = 2.0 * ((double)(xo_next(st) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
            r2 = u * u + v * v;
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = scale * sqrt(-2.0 * log(r2) / r2);
        z[i++] = u * f;
        if (i < N) z[i++] = v * f;
    }
}

static inline double Sg(double x, double sigma, double mu) { return 1.0 / (1.0 + exp(-(x - mu) * sigma)); }

/* CM [N,N] row-major; G, sigmaE [N]; noise NULL (philox) or [n1+n2+n3, N] already scaled by sqdtD.
 * Y   NULL or [nrec, 3, N]; Eonly NULL or [nrec, N] (E rows only); final NULL or [3, N]. */
int orc_wc_run(const orc_wc_params* p, const double* CM, int N, const double* G, const double* sigmaE,
               int64_t n1, int64_t n2, int64_t n3, int downsamp, int64_t nrec, const double* noise,
               uint64_t seed, uint64_t stream, double* Y, double* Eonly, double* final, int fast_rng) {
    double* buf = (double*)malloc(sizeof(double) * (size_t)N * 8);
    if (!buf) return -1;
    double *E = buf, *I = buf + N, *a = buf + 2 * N, *c = buf + 3 * N, *nz = buf + 4 * N;
    double *dE = buf + 5 * N, *dI = buf + 6 * N, *da = buf + 7 * N;
    for (int i = 0; i < N; ++i) { E[i] = p->E0; I[i] = p->I0; a[i] = p->a_ie_0; }
    const int64_t ns[3] = {n1, n2, n3};
    int64_t step = 0;
    uint64_t xst[4] = {seed ^ 0x9E3779B97F4A7C15ull, stream + 0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0x2545F4914F6CDD1Dull};
    for (int i = 0; i < 16; ++i) xo_next(xst);
    for (int ph = 0; ph < 3; ++ph) {
        const double tau_ip = p->tau_ip[ph];
        for (int64_t it = 0; it < ns[ph]; ++it, ++step) {
            if (ph == 2 && it % downsamp == 0) {
                int64_t r = it / downsamp;
                if (r < nrec) {
                    if (Y) {
                        memcpy(Y + (r * 3 + 0) * N, E, sizeof(double) * N);
                        memcpy(Y + (r * 3 + 1) * N, I, sizeof(double) * N);
                        memcpy(Y + (r * 3 + 2) * N, a, sizeof(double) * N);
                    }
                    if (Eonly) memcpy(Eonly + r * N, E, sizeof(double) * N);
                }
            }
            if (noise) {
                memcpy(nz, noise + step * N, sizeof(double) * N);
            } else if (fast_rng) {
                polar_normals(xst, N, p->sqdtD, nz);
            } else {
                orc_philox_normals(seed, stream, (uint32_t)step, N, nz);
                for (int i = 0; i < N; ++i) nz[i] *= p->sqdtD;
            }
            for (int i = 0; i < N; ++i) {
                const double* row = CM + (size_t)i * N;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     /* 4 partial sums: lets the compiler vectorise */
                int j = 0;
                for (; j + 3 < N; j += 4) {
                    a0 += row[j] * E[j]; a1 += row[j + 1] * E[j + 1];
                    a2 += row[j + 2] * E[j + 2]; a3 += row[j + 3] * E[j + 3];
                }
                for (; j < N; ++j) a0 += row[j] * E[j];
                c[i] = (a0 + a1) + (a2 + a3);
            }
            for (int i = 0; i < N; ++i) {
                dE[i] = (-E[i] + (1 - p->rE * E[i]) * Sg(p->a_ee * E[i] - a[i] * I[i] + G[i] * c[i] + p->P + nz[i], sigmaE[i], p->mu)) / p->tauE;
                dI[i] = (-I[i] + (1 - p->rI * I[i]) * Sg(p->a_ei * E[i] - p->a_ii * I[i], p->sigmaI, p->mu)) / p->tauI;
                da[i] = (I[i] * (E[i] - p->rhoE)) / tau_ip;
            }
            for (int i = 0; i < N; ++i) {
                E[i] += p->dtSim * dE[i];
                I[i] += p->dtSim * dI[i];
                a[i] += p->dtSim * da[i];
            }
        }
    }
    if (final) memcpy(final, buf, sizeof(double) * 3 * N);
    free(buf);
    return 0;
}

/* rE [T,N] -> out [T,N]; see oracle/bold_oracle.py:bold_sim */
int orc_bold_sim(const double* rE, int64_t T, int N, double dt, double* out) {
    const double kappa = 1.0 / 0.65, gamma = 1.0 / 0.41, tau = 0.98, alpha = 0.32, E0 = 0.4, V0 = 0.04, TE = 0.04;
    const double k1 = 4.3 * 40.3 * E0 * TE, k2 = 25.0 * E0 * TE, k3 = 1.0, ia = 1.0 / alpha;
    for (int n = 0; n < N; ++n) {
        double s = 0.1, f = 1.0, v = 1.0, q = 1.0;
        for (int64_t i = 0; i < T; ++i) {
            out[i * N + n] = V0 * (k1 * (1 - q) + k2 * (1 - q / v) + k3 * (1 - v));
            double x = rE[i * N + n];
            double va = pow(v, ia);
            double ds = x - kappa * s - gamma * (f - 1);
            double df = s;
            double dv = (f - va) / tau;
            double dq = (f * (1 - pow(1 - E0, 1 / f)) / E0 - q * va / v) / tau;
            s += dt * ds; f += dt * df; v += dt * dv; q += dt * dq;
        }
    }
    return 0;
}
