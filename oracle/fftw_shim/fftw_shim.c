/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * fp64 implementation of the six FFTW3 symbols the reference's host path
 * needs (see fftw3.h in this directory).  A length-n real transform is done
 * as a length-n/2 complex Stockham radix-4 FFT plus the usual even/odd
 * split/merge step.  Plans are immutable after creation; fftw_execute only
 * touches the in/out buffers given at plan time and a per-plan scratch area,
 * exactly like FFTW's "execute on the planned arrays" contract.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

struct fftw_shim_plan_s {
    int n;          /* real length                         */
    int h;          /* n/2 = complex length                */
    int inverse;    /* 0: r2c (exp -), 1: c2r (exp +)       */
    void *in;
    void *out;
    cpx *tw;        /* exp(sign*2*pi*i*k/h), k < h          */
    cpx *twn;       /* exp(sign*2*pi*i*k/n), k <= h         */
    cpx *buf0;
    cpx *buf1;
};

void *fftw_malloc(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
    return p;
}

void fftw_free(void *p) { free(p); }

static fftw_plan make_plan(int n, void *in, void *out, int inverse) {
    if (n < 8 || (n & (n - 1)) != 0) return NULL;
    fftw_plan p = (fftw_plan) malloc(sizeof(*p));
    p->n = n;
    p->h = n / 2;
    p->inverse = inverse;
    p->in = in;
    p->out = out;
    p->tw = (cpx *) fftw_malloc(sizeof(cpx) * p->h);
    p->twn = (cpx *) fftw_malloc(sizeof(cpx) * (p->h + 1));
    p->buf0 = (cpx *) fftw_malloc(sizeof(cpx) * (p->h + 1));
    p->buf1 = (cpx *) fftw_malloc(sizeof(cpx) * (p->h + 1));
    const double sgn = inverse ? 1.0 : -1.0;
    for (int k = 0; k < p->h; k++) {
        double a = 2.0 * M_PI * (double) k / (double) p->h;
        p->tw[k].re = cos(a);
        p->tw[k].im = sgn * sin(a);
    }
    for (int k = 0; k <= p->h; k++) {
        double a = 2.0 * M_PI * (double) k / (double) n;
        p->twn[k].re = cos(a);
        p->twn[k].im = sgn * sin(a);
    }
    return p;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags) {
    (void) flags;
    return make_plan(n, in, out, 0);
}

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags) {
    (void) flags;
    return make_plan(n, in, out, 1);
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    fftw_free(p->tw);
    fftw_free(p->twn);
    fftw_free(p->buf0);
    fftw_free(p->buf1);
    free(p);
}

/* Complex Stockham FFT of length h on x (result returned in the buffer whose
 * pointer is returned: either x or y).  tw[k] = exp(sign*2*pi*i*k/h); the
 * multiplication by +-i in the radix-4 butterfly takes its sign from tw. */
static cpx *cfft(int h, const cpx *tw, int inverse, cpx *x, cpx *y) {
    int n = h, s = 1;
    const double js = inverse ? 1.0 : -1.0; /* multiply by js*i */
    while (n >= 4) {
        const int m = n / 4;
        for (int p = 0; p < m; p++) {
            const cpx w1 = tw[(size_t) p * s];
            const cpx w2 = tw[(size_t) 2 * p * s];
            const cpx w3 = tw[(size_t) 3 * p * s];
            const cpx *x0 = x + (size_t) s * (p + 0 * m);
            const cpx *x1 = x + (size_t) s * (p + 1 * m);
            const cpx *x2 = x + (size_t) s * (p + 2 * m);
            const cpx *x3 = x + (size_t) s * (p + 3 * m);
            cpx *y0 = y + (size_t) s * (4 * p + 0);
            cpx *y1 = y + (size_t) s * (4 * p + 1);
            cpx *y2 = y + (size_t) s * (4 * p + 2);
            cpx *y3 = y + (size_t) s * (4 * p + 3);
            for (int q = 0; q < s; q++) {
                const double ar = x0[q].re, ai = x0[q].im;
                const double br = x1[q].re, bi = x1[q].im;
                const double cr = x2[q].re, ci = x2[q].im;
                const double dr = x3[q].re, di = x3[q].im;
                const double apcr = ar + cr, apci = ai + ci;
                const double amcr = ar - cr, amci = ai - ci;
                const double bpdr = br + dr, bpdi = bi + di;
                /* j*(b-d) with j = js*i */
                const double jr = -js * (bi - di), ji = js * (br - dr);
                y0[q].re = apcr + bpdr;
                y0[q].im = apci + bpdi;
                const double t1r = amcr + jr, t1i = amci + ji;
                const double t2r = apcr - bpdr, t2i = apci - bpdi;
                const double t3r = amcr - jr, t3i = amci - ji;
                y1[q].re = t1r * w1.re - t1i * w1.im;
                y1[q].im = t1r * w1.im + t1i * w1.re;
                y2[q].re = t2r * w2.re - t2i * w2.im;
                y2[q].im = t2r * w2.im + t2i * w2.re;
                y3[q].re = t3r * w3.re - t3i * w3.im;
                y3[q].im = t3r * w3.im + t3i * w3.re;
            }
        }
        n = m;
        s *= 4;
        cpx *t = x; x = y; y = t;
    }
    if (n == 2) {
        for (int q = 0; q < s; q++) {
            const cpx a = x[q], b = x[q + s];
            y[q].re = a.re + b.re;
            y[q].im = a.im + b.im;
            y[q + s].re = a.re - b.re;
            y[q + s].im = a.im - b.im;
        }
        cpx *t = x; x = y; y = t;
    }
    return x;
}

static void exec_r2c(const fftw_plan p, const double *in, cpx *out) {
    const int h = p->h;
    cpx *z = p->buf0;
    memcpy(z, in, sizeof(double) * (size_t) p->n);
    cpx *Z = cfft(h, p->tw, 0, z, p->buf1);
    /* X[k] = E[k] + w^k O[k];  E = (Z[k]+conj Z[h-k])/2,  O = (Z[k]-conj Z[h-k])/(2i) */
    for (int k = 0; k <= h; k++) {
        const cpx a = Z[k == h ? 0 : k];
        const cpx b = Z[(h - k) == h ? 0 : (h - k)];
        const double er = 0.5 * (a.re + b.re), ei = 0.5 * (a.im - b.im);
        const double dr = 0.5 * (a.re - b.re), di = 0.5 * (a.im + b.im);
        /* O = d/i = (di, -dr) */
        const double orr = di, oi = -dr;
        const cpx w = p->twn[k];
        out[k].re = er + (orr * w.re - oi * w.im);
        out[k].im = ei + (orr * w.im + oi * w.re);
    }
}

static void exec_c2r(const fftw_plan p, const cpx *in, double *out) {
    const int h = p->h;
    cpx *Z = p->buf0;
    for (int k = 0; k < h; k++) {
        const cpx a = in[k];
        const cpx b = in[h - k];
        const double er = a.re + b.re, ei = a.im - b.im;
        const double dr = a.re - b.re, di = a.im + b.im;
        const cpx w = p->twn[k];
        const double orr = dr * w.re - di * w.im;
        const double oi = dr * w.im + di * w.re;
        /* Z = E' + i*O' */
        Z[k].re = er - oi;
        Z[k].im = ei + orr;
    }
    cpx *z = cfft(h, p->tw, 1, Z, p->buf1);
    memcpy(out, z, sizeof(double) * (size_t) p->n);
}

void fftw_shim_execute_on(const fftw_plan p, void *in, void *out) {
    if (p->inverse) exec_c2r(p, (const cpx *) in, (double *) out);
    else exec_r2c(p, (const double *) in, (cpx *) out);
}

void fftw_execute(const fftw_plan p) { fftw_shim_execute_on(p, p->in, p->out); }

double *fftw_shim_cfft(int h, const double *tw, int inverse, double *x, double *y) {
    return (double *) cfft(h, (const cpx *) tw, inverse, (cpx *) x, (cpx *) y);
}
