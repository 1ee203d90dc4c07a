/*
 * qck_oracle.c - plain-C CPU restatement of the two heavy loops of the hot path.
 * TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY (see oracle/__init__.py): the
 * product never links or loads this.
 *
 *   oracle_knit_outer     QuasiDistr.merge over fragments without virtual gates
 *                         (third_party/qvm/qvm/quasi_distr.py:55-60 driven by
 *                         virtual_circuit.py:165-171,216-228): the XOR-key outer
 *                         product written densely,
 *                         out[y] = prod_f table_f[pext(y, mask_f)].
 *   oracle_knit_contract  closed form of the level loop (virtual_circuit.py:59-68
 *                         + virtual_gates.py knit rules):
 *                         out[y] = sum_l w(l) prod_f Q_f[l_f][pext(y, mask_f)].
 *   oracle_apply_1q / oracle_apply_cx / oracle_apply_cz / oracle_apply_2q
 *                         statevector gate application (the part of the path the
 *                         reference delegates to qiskit-aer 0.13.0, third-party,
 *                         not vendored; call site run.py:42).
 *
 * OpenMP threads = the host cores; every loop is the textbook form, no blocking
 * or SIMD tricks - it is the baseline, not the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t pext64(uint64_t y, uint64_t mask) {
    uint64_t out = 0;
    int j = 0;
    while (mask) {
        uint64_t low = mask & (~mask + 1);
        if (y & low) out |= (1ull << j);
        ++j;
        mask ^= low;
    }
    return out;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU baseline sets its thread count itself */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* returns sum of the produced entries through *sum_out and their minimum through *min_out */
void oracle_knit_outer(int n_frag, const double* const* tables, const uint64_t* masks, uint64_t y_begin,
                       uint64_t y_end, double* out, double* sum_out, double* min_out) {
    double sum = 0.0, mn = INFINITY;
#pragma omp parallel for schedule(static) reduction(+ : sum) reduction(min : mn)
    for (int64_t y = (int64_t)y_begin; y < (int64_t)y_end; ++y) {
        double v = 1.0;
        for (int f = 0; f < n_frag; ++f) v *= tables[f][pext64((uint64_t)y, masks[f])];
        if (out) out[(uint64_t)y - y_begin] = v;
        sum += v;
        if (v < mn) mn = v;
    }
    if (sum_out) *sum_out = sum;
    if (min_out) *min_out = mn;
}

/* weights[l], rows[f * n_labels + l] prepared by the caller (numpy) */
void oracle_knit_contract(int n_frag, const double* const* tables, const uint64_t* masks,
                          const int64_t* row_strides, int n_out_bits, int64_t n_labels, const double* weights,
                          const int32_t* rows, double* out) {
    const int64_t n = (int64_t)1 << n_out_bits;
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < n; ++y) {
        const double* base[16];
        for (int f = 0; f < n_frag; ++f) base[f] = tables[f] + pext64((uint64_t)y, masks[f]);
        double acc = 0.0;
        for (int64_t l = 0; l < n_labels; ++l) {
            double t = weights[l];
            for (int f = 0; f < n_frag; ++f) t *= base[f][(int64_t)rows[f * n_labels + l] * row_strides[f]];
            acc += t;
        }
        out[y] = acc;
    }
}

/* state: interleaved re/im, 2^n amplitudes; m: row-major 2x2 complex (8 doubles) */
void oracle_apply_1q(double* state, int n, int q, const double* m) {
    const int64_t half = (int64_t)1 << (n - 1);
    const int64_t bit = (int64_t)1 << q;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < half; ++p) {
        int64_t i0 = ((p >> q) << (q + 1)) | (p & (bit - 1)), i1 = i0 | bit;
        double ar = state[2 * i0], ai = state[2 * i0 + 1], br = state[2 * i1], bi = state[2 * i1 + 1];
        state[2 * i0] = m[0] * ar - m[1] * ai + m[2] * br - m[3] * bi;
        state[2 * i0 + 1] = m[0] * ai + m[1] * ar + m[2] * bi + m[3] * br;
        state[2 * i1] = m[4] * ar - m[5] * ai + m[6] * br - m[7] * bi;
        state[2 * i1 + 1] = m[4] * ai + m[5] * ar + m[6] * bi + m[7] * br;
    }
}

void oracle_apply_cx(double* state, int n, int c, int t) {
    const int64_t total = (int64_t)1 << n;
    const int64_t cb = (int64_t)1 << c, tb = (int64_t)1 << t;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) {
        if ((i & cb) && !(i & tb)) {
            int64_t j = i | tb;
            double r = state[2 * i], im = state[2 * i + 1];
            state[2 * i] = state[2 * j];
            state[2 * i + 1] = state[2 * j + 1];
            state[2 * j] = r;
            state[2 * j + 1] = im;
        }
    }
}

void oracle_apply_cz(double* state, int n, int a, int b) {
    const int64_t total = (int64_t)1 << n;
    const int64_t m = ((int64_t)1 << a) | ((int64_t)1 << b);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) {
        if ((i & m) == m) {
            state[2 * i] = -state[2 * i];
            state[2 * i + 1] = -state[2 * i + 1];
        }
    }
}

/* m: row-major 4x4 complex, index = bit(q0) + 2 bit(q1) */
void oracle_apply_2q(double* state, int n, int q0, int q1, const double* m) {
    const int64_t total = (int64_t)1 << n;
    const int64_t b0 = (int64_t)1 << q0, b1 = (int64_t)1 << q1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) {
        if (i & (b0 | b1)) continue;
        const int64_t idx[4] = {i, i | b0, i | b1, i | b0 | b1};
        double ar[4], ai[4];
        for (int k = 0; k < 4; ++k) {
            ar[k] = state[2 * idx[k]];
            ai[k] = state[2 * idx[k] + 1];
        }
        for (int r = 0; r < 4; ++r) {
            double sr = 0.0, si = 0.0;
            for (int k = 0; k < 4; ++k) {
                const double mr = m[2 * (4 * r + k)], mi = m[2 * (4 * r + k) + 1];
                sr += mr * ar[k] - mi * ai[k];
                si += mr * ai[k] + mi * ar[k];
            }
            state[2 * idx[r]] = sr;
            state[2 * idx[r] + 1] = si;
        }
    }
}

/* prob[i] = |amp_i|^2 */
void oracle_probabilities(const double* state, int n, double* prob) {
    const int64_t total = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) prob[i] = state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
}
