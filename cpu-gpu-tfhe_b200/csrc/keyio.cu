// Key and ciphertext FILE FORMATS of stock TFHE clients (host code only).
//
// The reference's client -> cloud hand-off goes through three files (cpu/main.cpp:26-71 writes
// secret.key / cloud.key / cloud.data, cpu/cloud.cpp:138-161 reads cloud.key / cloud.data and
// writes answer.data) in the serialisation of gpuParallel/tfhe_io.cu:
//   * parameter sections are TEXT blocks  -----BEGIN X----- / "name: value" lines in std::map
//     (alphabetical) order, doubles printed with %.8lf / -----END X-----
//     (tfhe_generic_streams.cu:43-53, 107-166);
//   * payloads are raw little-endian binary: an int32 type id, then the words
//     (ids: tfhe_generic_streams.h:15-30).
// cloud key  = GATEBOOTSPARAMS, LWEPARAMS, TLWEPARAMS, TGSWPARAMS, LWEKSPARAMS,
//              [200][double variance] ks[N][t][base][n+1],
//              [201][double variance] bk[n][kpl][k+1][N]           (tfhe_io.cu:757-814, 883-970, 1099-1103)
// secret key = the same, then [43] lwe_key[n], [169] tlwe_key[k][N]   (:1160-1166)
// ciphertext = [42] a[n] b [double variance]                          (:90-108)
// The payload orders are exactly this library's flat key formats (tfhe_b200.h), so reading a
// key file is a header parse plus two freads straight into the arrays tfhe_b200_load_keys takes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/tfhe_compat.h"

namespace {

constexpr int32_t kUidLweSample = 42, kUidLweKey = 43, kUidTgswKey = 169, kUidKs = 200, kUidBk = 201;

thread_local char g_ioerr[256] = "";

int io_fail(const char *fmt, const char *arg = "") {
    snprintf(g_ioerr, sizeof(g_ioerr), fmt, arg);
    fprintf(stderr, "tfhe_b200 io: %s\n", g_ioerr);
    return 1;
}

typedef std::map<std::string, std::string> Props;

std::string fmt_double(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.8lf", v);  // setProperty_double, tfhe_generic_streams.cu:43-47
    return buf;
}

std::string fmt_long(long v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%ld", v);
    return buf;
}

void put_section(FILE *f, const char *title, const Props &p) {
    fprintf(f, "-----BEGIN %s-----\n", title);
    for (const auto &kv : p) fprintf(f, "%s: %s\n", kv.first.c_str(), kv.second.c_str());
    fprintf(f, "-----END %s-----\n", title);
}

bool get_line(FILE *f, std::string &line) {  // CIstream::getLine, tfhe_generic_streams.cu:67-74
    line.clear();
    int c = fgetc(f);
    if (c == EOF) return false;
    for (; c != EOF; c = fgetc(f)) {
        if (c == '\r') continue;
        if (c == '\n') return true;
        line.push_back((char) c);
    }
    return true;
}

// new_TextModeProperties_fromIstream, tfhe_generic_streams.cu:118-151
bool get_section(FILE *f, const char *want, Props &p) {
    p.clear();
    std::string line, title, end;
    bool started = false;
    while (get_line(f, line)) {
        const size_t n = line.size();
        if (n >= 16 && line.compare(0, 11, "-----BEGIN ") == 0 && line.compare(n - 5, 5, "-----") == 0) {
            title = line.substr(11, n - 16);
            end = "-----END " + title + "-----";
            started = true;
            continue;
        }
        if (!started) continue;  // anything before the body is ignored
        if (line == end) return title == want;
        const size_t pos = line.find(": ");
        if (pos == std::string::npos) continue;
        p[line.substr(0, pos)] = line.substr(pos + 2);
    }
    return false;
}

bool has(const Props &p, const char *k) { return p.find(k) != p.end(); }

struct Header {
    tfhe_b200_params p;
    double alphas[4];  // lwe alpha_min, alpha_max, tlwe alpha_min, alpha_max
};

void write_header(FILE *f, const tfhe_b200_params &p, const double *alphas) {
    // write_tfheGateBootstrappingParameters (tfhe_io.cu:1031-1035) + LWEKSPARAMS (:731-739)
    put_section(f, "GATEBOOTSPARAMS", {{"ks_basebit", fmt_long(p.ks_basebit)}, {"ks_t", fmt_long(p.ks_t)}});
    put_section(f, "LWEPARAMS",
                {{"alpha_max", fmt_double(alphas[1])}, {"alpha_min", fmt_double(alphas[0])}, {"n", fmt_long(p.n)}});
    put_section(f, "TLWEPARAMS", {{"N", fmt_long(p.N)},
                                  {"alpha_max", fmt_double(alphas[3])},
                                  {"alpha_min", fmt_double(alphas[2])},
                                  {"k", fmt_long(p.k)}});
    put_section(f, "TGSWPARAMS", {{"Bgbit", fmt_long(p.Bgbit)}, {"l", fmt_long(p.l)}});
    put_section(f, "LWEKSPARAMS",
                {{"basebit", fmt_long(p.ks_basebit)}, {"n", fmt_long((long) p.N * p.k)}, {"t", fmt_long(p.ks_t)}});
}

int read_header(FILE *f, Header &h) {
    Props s;
    if (!get_section(f, "GATEBOOTSPARAMS", s) || !has(s, "ks_t") || !has(s, "ks_basebit"))
        return io_fail("missing GATEBOOTSPARAMS section");
    h.p.ks_t = (int32_t) strtol(s["ks_t"].c_str(), nullptr, 10);
    h.p.ks_basebit = (int32_t) strtod(s["ks_basebit"].c_str(), nullptr);  // read as a double, tfhe_io.cu:1027
    if (!get_section(f, "LWEPARAMS", s) || !has(s, "n")) return io_fail("missing LWEPARAMS section");
    h.p.n = (int32_t) strtol(s["n"].c_str(), nullptr, 10);
    h.alphas[0] = strtod(s["alpha_min"].c_str(), nullptr);
    h.alphas[1] = strtod(s["alpha_max"].c_str(), nullptr);
    if (!get_section(f, "TLWEPARAMS", s) || !has(s, "N") || !has(s, "k")) return io_fail("missing TLWEPARAMS section");
    h.p.N = (int32_t) strtol(s["N"].c_str(), nullptr, 10);
    h.p.k = (int32_t) strtol(s["k"].c_str(), nullptr, 10);
    h.alphas[2] = strtod(s["alpha_min"].c_str(), nullptr);
    h.alphas[3] = strtod(s["alpha_max"].c_str(), nullptr);
    if (!get_section(f, "TGSWPARAMS", s) || !has(s, "l") || !has(s, "Bgbit")) return io_fail("missing TGSWPARAMS section");
    h.p.l = (int32_t) strtol(s["l"].c_str(), nullptr, 10);
    h.p.Bgbit = (int32_t) strtol(s["Bgbit"].c_str(), nullptr, 10);
    if (!get_section(f, "LWEKSPARAMS", s) || !has(s, "n") || !has(s, "t") || !has(s, "basebit"))
        return io_fail("missing LWEKSPARAMS section");
    if (strtol(s["n"].c_str(), nullptr, 10) != (long) h.p.N * h.p.k)
        return io_fail("wrong dimension in bootstrapping key");  // tfhe_io.cu:964-965
    if (strtol(s["t"].c_str(), nullptr, 10) != h.p.ks_t || strtol(s["basebit"].c_str(), nullptr, 10) != h.p.ks_basebit)
        return io_fail("key-switch parameters disagree with the gate parameters");
    if (h.p.n < 1 || h.p.N < 1 || h.p.k < 1 || h.p.l < 1 || h.p.ks_t < 1 || h.p.ks_basebit < 1 || h.p.ks_basebit > 8)
        return io_fail("implausible parameters");
    return 0;
}

int read_block(FILE *f, int32_t uid, double *variance, int32_t *dst, size_t words, const char *what) {
    int32_t got = -1;
    if (fread(&got, sizeof(got), 1, f) != 1 || got != uid) return io_fail("bad type id in the %s section", what);
    if (variance && fread(variance, sizeof(double), 1, f) != 1) return io_fail("truncated %s section", what);
    if (dst) {
        if (fread(dst, sizeof(int32_t), words, f) != words) return io_fail("truncated %s section", what);
    } else if (fseek(f, (long) (words * sizeof(int32_t)), SEEK_CUR) != 0) {
        return io_fail("truncated %s section", what);
    }
    return 0;
}

void write_block(FILE *f, int32_t uid, const double *variance, const int32_t *src, size_t words) {
    fwrite(&uid, sizeof(uid), 1, f);
    if (variance) fwrite(variance, sizeof(double), 1, f);
    fwrite(src, sizeof(int32_t), words, f);
}

int read_key_stream(FILE *f, tfhe_b200_params *p, double *alphas, double *variances, int32_t *bk, int32_t *ks,
                    int32_t *lwe_key, int32_t *tlwe_key, bool secret) {
    Header h;
    if (read_header(f, h)) return 1;
    if (p) *p = h.p;
    if (alphas) memcpy(alphas, h.alphas, sizeof(h.alphas));
    if (!bk && !ks && !lwe_key && !tlwe_key && !variances) return 0;  // header only
    double var_ks = 0, var_bk = 0;
    if (read_block(f, kUidKs, &var_ks, ks, tfhe_b200_ks_words(&h.p), "key-switch key")) return 1;
    if (read_block(f, kUidBk, &var_bk, bk, tfhe_b200_bk_words(&h.p), "bootstrapping key")) return 1;
    if (variances) {
        variances[0] = var_bk;
        variances[1] = var_ks;
    }
    if (secret) {
        if (read_block(f, kUidLweKey, nullptr, lwe_key, (size_t) h.p.n, "LWE key")) return 1;
        if (read_block(f, kUidTgswKey, nullptr, tlwe_key, (size_t) h.p.k * h.p.N, "TGSW key")) return 1;
    }
    return 0;
}

void write_key_stream(FILE *f, const tfhe_b200_params *p, const double *alphas, const double *variances,
                      const int32_t *bk, const int32_t *ks, const int32_t *lwe_key, const int32_t *tlwe_key) {
    const double zero4[4] = {0, 0, 0, 0};
    write_header(f, *p, alphas ? alphas : zero4);
    const double var_bk = variances ? variances[0] : 0.0, var_ks = variances ? variances[1] : 0.0;
    write_block(f, kUidKs, &var_ks, ks, tfhe_b200_ks_words(p));
    write_block(f, kUidBk, &var_bk, bk, tfhe_b200_bk_words(p));
    if (lwe_key) write_block(f, kUidLweKey, nullptr, lwe_key, (size_t) p->n);
    if (tlwe_key) write_block(f, kUidTgswKey, nullptr, tlwe_key, (size_t) p->k * p->N);
}

int read_samples(FILE *f, int n, int32_t *samples, double *variances, int count) {
    for (int i = 0; i < count; i++) {
        int32_t uid = -1;
        double var = 0;
        if (fread(&uid, sizeof(uid), 1, f) != 1 || uid != kUidLweSample) return io_fail("bad type id in ciphertext");
        if (fread(samples + (size_t) i * (n + 1), sizeof(int32_t), (size_t) n + 1, f) != (size_t) n + 1 ||
            fread(&var, sizeof(var), 1, f) != 1)
            return io_fail("truncated ciphertext");
        if (variances) variances[i] = var;
    }
    return 0;
}

void write_samples(FILE *f, int n, const int32_t *samples, const double *variances, int count) {
    for (int i = 0; i < count; i++) {
        const double var = variances ? variances[i] : 0.0;
        fwrite(&kUidLweSample, sizeof(int32_t), 1, f);
        fwrite(samples + (size_t) i * (n + 1), sizeof(int32_t), (size_t) n + 1, f);
        fwrite(&var, sizeof(var), 1, f);
    }
}

}  // namespace

extern "C" {

// ---- flat interface ----------------------------------------------------------------------

int tfhe_b200_file_read_cloud_key(const char *path, tfhe_b200_params *p, double *alphas4, double *variances2,
                                  int32_t *bk_coef, int32_t *ks) {
    FILE *f = fopen(path, "rb");
    if (!f) return io_fail("cannot open %s", path);
    const int rc = read_key_stream(f, p, alphas4, variances2, bk_coef, ks, nullptr, nullptr, false);
    fclose(f);
    return rc;
}

int tfhe_b200_file_write_cloud_key(const char *path, const tfhe_b200_params *p, const double *alphas4,
                                   const double *variances2, const int32_t *bk_coef, const int32_t *ks) {
    if (!p || !bk_coef || !ks) return io_fail("null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return io_fail("cannot create %s", path);
    write_key_stream(f, p, alphas4, variances2, bk_coef, ks, nullptr, nullptr);
    return fclose(f) ? io_fail("write to %s failed", path) : 0;
}

int tfhe_b200_file_read_secret_key(const char *path, tfhe_b200_params *p, double *alphas4, double *variances2,
                                   int32_t *bk_coef, int32_t *ks, int32_t *lwe_key, int32_t *tlwe_key) {
    FILE *f = fopen(path, "rb");
    if (!f) return io_fail("cannot open %s", path);
    const int rc = read_key_stream(f, p, alphas4, variances2, bk_coef, ks, lwe_key, tlwe_key, true);
    fclose(f);
    return rc;
}

int tfhe_b200_file_write_secret_key(const char *path, const tfhe_b200_params *p, const double *alphas4,
                                    const double *variances2, const int32_t *bk_coef, const int32_t *ks,
                                    const int32_t *lwe_key, const int32_t *tlwe_key) {
    if (!p || !bk_coef || !ks || !lwe_key || !tlwe_key) return io_fail("null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return io_fail("cannot create %s", path);
    write_key_stream(f, p, alphas4, variances2, bk_coef, ks, lwe_key, tlwe_key);
    return fclose(f) ? io_fail("write to %s failed", path) : 0;
}

// number of ciphertext records in a file of samples of dimension n (-1: not a whole number)
long tfhe_b200_file_count_ciphertexts(const char *path, int n) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fclose(f);
    const long rec = 4 + 4L * (n + 1) + 8;
    return bytes % rec ? -1 : bytes / rec;
}

int tfhe_b200_file_read_ciphertexts(const char *path, int n, int32_t *samples, double *variances, int count) {
    if (!samples || count < 0) return io_fail("bad argument");
    FILE *f = fopen(path, "rb");
    if (!f) return io_fail("cannot open %s", path);
    const int rc = read_samples(f, n, samples, variances, count);
    fclose(f);
    return rc;
}

int tfhe_b200_file_write_ciphertexts(const char *path, int n, const int32_t *samples, const double *variances,
                                     int count, int append) {
    if (!samples || count < 0) return io_fail("bad argument");
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return io_fail("cannot create %s", path);
    write_samples(f, n, samples, variances, count);
    return fclose(f) ? io_fail("write to %s failed", path) : 0;
}

const char *tfhe_b200_file_last_error(void) { return g_ioerr; }

// ---- the reference's own entry points (tfhe_io.h), on the reference's structs ----------------
// A cloud key set read here carries the coefficient-domain bootstrapping key (bk) and the
// key-switch key; bkFFT stays NULL: the gate functions of this library convert on the GPU when
// the key set is first used (compat.cu, ctx_for_cloud).

struct KeySetBlock {  // one allocation owning everything a key set read from a file points to
    TFheGateBootstrappingCloudKeySet cloud;
    TFheGateBootstrappingParameterSet params;
    LweParams lwe;
    TLweParams tlwe;
    TGswParams tgsw;
    LweBootstrappingKey bk;
    LweKeySwitchKey ks;
    std::vector<Torus32> h;
    std::vector<int32_t> bk_words, ks_words;
    std::vector<TGswSample> tgsw_samples;
    std::vector<TLweSample> tlwe_samples;
    std::vector<TLweSample *> bloc;
    std::vector<TorusPolynomial> polys;
    std::vector<LweSample> ks0;
    std::vector<LweSample *> ks1;
    std::vector<LweSample **> ks2;
};

static void fill_params(KeySetBlock *B, const tfhe_b200_params &p, const double *alphas) {
    B->lwe = LweParams{p.n, alphas[0], alphas[1]};
    B->tlwe.N = p.N;
    B->tlwe.k = p.k;
    B->tlwe.alpha_min = alphas[2];
    B->tlwe.alpha_max = alphas[3];
    B->tlwe.extracted_lweparams = LweParams{p.N * p.k, alphas[2], alphas[3]};  // tlwe.cu: TLweParams ctor
    // TGswParams ctor, tgsw.cu:7-29
    B->tgsw.l = p.l;
    B->tgsw.Bgbit = p.Bgbit;
    B->tgsw.Bg = 1 << p.Bgbit;
    B->tgsw.halfBg = B->tgsw.Bg / 2;
    B->tgsw.maskMod = (uint32_t) B->tgsw.Bg - 1;
    B->tgsw.tlwe_params = &B->tlwe;
    B->tgsw.kpl = (p.k + 1) * p.l;
    B->h.resize(p.l);
    uint32_t off = 0;
    for (int i = 0; i < p.l; i++) {
        const int kk = 32 - (i + 1) * p.Bgbit;
        B->h[i] = (Torus32) (1u << kk);
        off += 1u << kk;
    }
    B->tgsw.h = B->h.data();
    B->tgsw.offset = off * (uint32_t) B->tgsw.halfBg;
    B->params.ks_t = p.ks_t;
    B->params.ks_basebit = p.ks_basebit;
    B->params.in_out_params = &B->lwe;
    B->params.tgsw_params = &B->tgsw;
}

TFheGateBootstrappingCloudKeySet *new_tfheGateBootstrappingCloudKeySet_fromFile(FILE *F) {
    KeySetBlock *B = new KeySetBlock();
    Header h;
    if (read_header(F, h)) abort();  // the reference aborts on malformed input (tfhe_io.cu:52, 964)
    const tfhe_b200_params &p = h.p;
    fill_params(B, p, h.alphas);
    B->bk_words.resize(tfhe_b200_bk_words(&p));
    B->ks_words.resize(tfhe_b200_ks_words(&p));
    double var_ks = 0, var_bk = 0;
    if (read_block(F, kUidKs, &var_ks, B->ks_words.data(), B->ks_words.size(), "key-switch key")) abort();
    if (read_block(F, kUidBk, &var_bk, B->bk_words.data(), B->bk_words.size(), "bootstrapping key")) abort();
    const int n = p.n, N = p.N, k = p.k, kpl = B->tgsw.kpl, t = p.ks_t, base = 1 << p.ks_basebit;
    // bootstrapping key: n TGSW samples of kpl TLWE samples of k+1 polynomials over the flat words
    B->polys.resize((size_t) n * kpl * (k + 1));
    B->tlwe_samples.resize((size_t) n * kpl);
    B->bloc.resize((size_t) n * (k + 1));
    B->tgsw_samples.resize(n);
    for (int i = 0; i < n; i++) {
        for (int r = 0; r < kpl; r++) {
            TLweSample &s = B->tlwe_samples[(size_t) i * kpl + r];
            s.a = &B->polys[((size_t) i * kpl + r) * (k + 1)];
            for (int j = 0; j <= k; j++) {
                s.a[j].N = N;
                s.a[j].coefsT = B->bk_words.data() + (((size_t) i * kpl + r) * (k + 1) + j) * N;
            }
            s.b = s.a + k;
            s.current_variance = var_bk;
            s.k = k;
        }
        TGswSample &g = B->tgsw_samples[i];
        g.all_sample = &B->tlwe_samples[(size_t) i * kpl];
        g.bloc_sample = &B->bloc[(size_t) i * (k + 1)];
        for (int j = 0; j <= k; j++) g.bloc_sample[j] = g.all_sample + j * p.l;
        g.k = k;
        g.l = p.l;
    }
    // key-switch key: ks[i][j][h] over the flat words (a then b per sample, b copied out)
    B->ks0.resize((size_t) N * k * t * base);
    B->ks1.resize((size_t) N * k * t);
    B->ks2.resize((size_t) N * k);
    for (size_t s = 0; s < B->ks0.size(); s++) {
        int32_t *w = B->ks_words.data() + s * (n + 1);
        B->ks0[s].a = w;
        B->ks0[s].b = w[n];
        B->ks0[s].current_variance = var_ks;
    }
    for (size_t s = 0; s < B->ks1.size(); s++) B->ks1[s] = &B->ks0[s * base];
    for (size_t s = 0; s < B->ks2.size(); s++) B->ks2[s] = &B->ks1[s * t];
    B->ks = LweKeySwitchKey{N * k, t, p.ks_basebit, base, &B->lwe, B->ks0.data(), B->ks1.data(), B->ks2.data()};
    B->bk = LweBootstrappingKey{&B->lwe, &B->tgsw, &B->tlwe, &B->tlwe.extracted_lweparams, B->tgsw_samples.data(),
                                &B->ks};
    B->cloud = TFheGateBootstrappingCloudKeySet{&B->params, &B->bk, nullptr};
    return &B->cloud;
}

// for key sets made by new_tfheGateBootstrappingCloudKeySet_fromFile of THIS library
void tfhe_b200_delete_cloud_keyset_fromFile(TFheGateBootstrappingCloudKeySet *ks) {
    if (!ks) return;
    tfhe_b200_keys_free(ks);
    delete reinterpret_cast<KeySetBlock *>(ks);  // cloud is the first member
}

void export_tfheGateBootstrappingCloudKeySet_toFile(FILE *F, const TFheGateBootstrappingCloudKeySet *keyset) {
    const TFheGateBootstrappingParameterSet *gp = keyset->params;
    const TLweParams *tp = gp->tgsw_params->tlwe_params;
    tfhe_b200_params p = {gp->in_out_params->n, tp->N, tp->k, gp->tgsw_params->l, gp->tgsw_params->Bgbit,
                          gp->ks_t, gp->ks_basebit};
    const double alphas[4] = {gp->in_out_params->alpha_min, gp->in_out_params->alpha_max, tp->alpha_min, tp->alpha_max};
    write_header(F, p, alphas);
    const LweBootstrappingKey *bk = keyset->bk;
    const LweKeySwitchKey *ks = bk->ks;
    const int n = p.n, N = p.N, k = p.k, kpl = gp->tgsw_params->kpl, base = ks->base;
    double var = -1;  // the maximum variance is written once (tfhe_io.cu:765-777)
    for (int i = 0; i < ks->n; i++)
        for (int j = 0; j < ks->t; j++)
            for (int h = 0; h < base; h++)
                if (ks->ks[i][j][h].current_variance > var) var = ks->ks[i][j][h].current_variance;
    fwrite(&kUidKs, sizeof(int32_t), 1, F);
    fwrite(&var, sizeof(double), 1, F);
    for (int i = 0; i < ks->n; i++)
        for (int j = 0; j < ks->t; j++)
            for (int h = 0; h < base; h++) {
                fwrite(ks->ks[i][j][h].a, sizeof(Torus32), (size_t) n, F);
                fwrite(&ks->ks[i][j][h].b, sizeof(Torus32), 1, F);
            }
    var = -1;
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            if (bk->bk[i].all_sample[r].current_variance > var) var = bk->bk[i].all_sample[r].current_variance;
    fwrite(&kUidBk, sizeof(int32_t), 1, F);
    fwrite(&var, sizeof(double), 1, F);
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++) fwrite(bk->bk[i].all_sample[r].a[j].coefsT, sizeof(Torus32), (size_t) N, F);
}

void export_gate_bootstrapping_ciphertext_toFile(FILE *F, const LweSample *sample,
                                                 const TFheGateBootstrappingParameterSet *params) {
    const int n = params->in_out_params->n;
    fwrite(&kUidLweSample, sizeof(int32_t), 1, F);
    fwrite(sample->a, sizeof(Torus32), (size_t) n, F);
    fwrite(&sample->b, sizeof(Torus32), 1, F);
    fwrite(&sample->current_variance, sizeof(double), 1, F);
}

void import_gate_bootstrapping_ciphertext_fromFile(FILE *F, LweSample *sample,
                                                   const TFheGateBootstrappingParameterSet *params) {
    const int n = params->in_out_params->n;
    int32_t uid = -1;
    if (fread(&uid, sizeof(uid), 1, F) != 1 || uid != kUidLweSample) abort();  // tfhe_io.cu:94
    if (fread(sample->a, sizeof(Torus32), (size_t) n, F) != (size_t) n || fread(&sample->b, sizeof(Torus32), 1, F) != 1 ||
        fread(&sample->current_variance, sizeof(double), 1, F) != 1)
        abort();
}

// new_gate_bootstrapping_ciphertext_array / delete_... (tfhe_gate_bootstrapping.cu:93-108): the
// caller-allocated result buffers of the classic API
LweSample *new_gate_bootstrapping_ciphertext_array(int nbelems, const TFheGateBootstrappingParameterSet *params) {
    const int n = params->in_out_params->n;
    LweSample *arr = (LweSample *) malloc(sizeof(LweSample) * (size_t) (nbelems > 0 ? nbelems : 1));
    Torus32 *words = (Torus32 *) calloc((size_t) (nbelems > 0 ? nbelems : 1) * n, sizeof(Torus32));
    if (!arr || !words) abort();
    for (int i = 0; i < nbelems; i++) {
        arr[i].a = words + (size_t) i * n;
        arr[i].b = 0;
        arr[i].current_variance = 0.;
    }
    if (nbelems <= 0) arr[0].a = words;
    return arr;
}

LweSample *new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet *params) {
    return new_gate_bootstrapping_ciphertext_array(1, params);
}

void delete_gate_bootstrapping_ciphertext_array(int nbelems, LweSample *samples) {
    (void) nbelems;
    if (!samples) return;
    free(samples[0].a);
    free(samples);
}

void delete_gate_bootstrapping_ciphertext(LweSample *sample) { delete_gate_bootstrapping_ciphertext_array(1, sample); }

}  // extern "C"
