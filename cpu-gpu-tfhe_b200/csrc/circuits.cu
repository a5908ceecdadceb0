// Cipher-level arithmetic circuits as device-resident gate schedules.
//
// The reference builds these from its batched gates with per-level buffer shuffling
// (taskLevelParallelAdd main.cu:619, taskLevelParallelAdd_bitwise :821,
// taskLevelParallelAdd_bitwise_vector_coalInput :1138, multiplyLweSamples :1483,
// BOOTS_vectorMultiplication :1746, BOOTS_matrixMultiplication :2342; CPU: Cipher::addBits /
// operator+ / operator* Cipher.cu:334-378, 83-108).  They add no arithmetic of their own:
// a circuit is a list of LEVELS, every level one batch of independent bootstrapped gates.
//
// Here a circuit is compiled once into a PLAN: all ciphertexts of the circuit live as rows of
// one device workspace, every level is one blind-rotate launch + one key-switch launch whose
// operand / result rows are given by index tables uploaded at plan creation (the shifts and
// re-layouts between levels, which the reference does with copy kernels and host memcpy, are
// just different indices).  Running a plan issues no host synchronisation.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tfhe_b200.h"

namespace {

struct Op {
    int gate, count, off_a, off_b, off_out;
};

struct Level {
    Op ops[4];
    int nops = 0;
};

}  // namespace

struct tfhe_b200_circuit {
    tfhe_b200_ctx *ctx = nullptr;
    int words = 0;
    int nrows = 0;
    std::vector<int> in_row0, in_rows;  // operand o occupies rows [in_row0[o], +in_rows[o])
    int out_row0 = 0, out_rows = 0;
    int row_zero = -1;                  // trivial encryption of 0 (bootsCONSTANT, boot-gates.cu:263)
    std::vector<Level> levels;
    std::vector<int32_t> h_idx;
    int32_t *d_idx = nullptr;
    int32_t *d_ws = nullptr;
    long long n_gates = 0;
};

namespace {

// ---- plan builder -----------------------------------------------------------------------

struct Builder {
    tfhe_b200_circuit *c;
    // gates of the level under construction, grouped by gate type
    std::vector<int> ga[TFHE_B200_NUM_GATES], gb[TFHE_B200_NUM_GATES], go[TFHE_B200_NUM_GATES];

    explicit Builder(tfhe_b200_circuit *circ) : c(circ) {}

    int alloc(int rows = 1) {
        const int r = c->nrows;
        c->nrows += rows;
        return r;
    }

    int operand(int rows) {
        const int r = alloc(rows);
        c->in_row0.push_back(r);
        c->in_rows.push_back(rows);
        return r;
    }

    // schedule out = gate(a, b) in the current level
    void gate(int g, int a, int b, int out) {
        ga[g].push_back(a);
        gb[g].push_back(b);
        go[g].push_back(out);
    }

    // close the level: one bootstrap batch per (up to) 4 gate types
    void end_level() {
        Level lv;
        for (int g = 0; g < TFHE_B200_NUM_GATES; g++) {
            if (ga[g].empty()) continue;
            if (lv.nops == 4) {
                c->levels.push_back(lv);
                lv = Level();
            }
            Op &op = lv.ops[lv.nops++];
            op.gate = g;
            op.count = (int) ga[g].size();
            op.off_a = (int) c->h_idx.size();
            c->h_idx.insert(c->h_idx.end(), ga[g].begin(), ga[g].end());
            op.off_b = (int) c->h_idx.size();
            c->h_idx.insert(c->h_idx.end(), gb[g].begin(), gb[g].end());
            op.off_out = (int) c->h_idx.size();
            c->h_idx.insert(c->h_idx.end(), go[g].begin(), go[g].end());
            c->n_gates += op.count;
            ga[g].clear();
            gb[g].clear();
            go[g].clear();
        }
        if (lv.nops) c->levels.push_back(lv);
    }

    // Ripple-carry addition of `m` pairs of nbits-bit numbers in lock-step
    // (taskLevelParallelAdd_bitwise[_vector_coalInput], main.cu:821-890 / 1138-1302; the
    // 5-gate full adder of Cipher::addBits, Cipher.cu:367-378, as 3 levels of 2/1/2 gates).
    // a[i], b[i], out[i]: rows of bit 0 of number i (bits are consecutive rows; -1 marks an
    // absent operand handled by the caller).  Result truncated to nbits.
    void ripple_add(const std::vector<int> &a, const std::vector<int> &b, const std::vector<int> &out, int nbits) {
        const int m = (int) a.size();
        const int carry = alloc(m), t0 = alloc(m), t1 = alloc(m);
        for (int i = 0; i < m; i++) {  // bit 0: (carry, sum) = (AND, XOR), bootsANDXOR main.cu:849
            if (nbits > 1) gate(TFHE_B200_AND, a[i], b[i], carry + i);
            gate(TFHE_B200_XOR, a[i], b[i], out[i]);
        }
        end_level();
        for (int bit = 1; bit < nbits; bit++) {
            for (int i = 0; i < m; i++) {  // t0 = a ^ c, t1 = b ^ c  (bootsXORXOR main.cu:869)
                gate(TFHE_B200_XOR, a[i] + bit, carry + i, t0 + i);
                gate(TFHE_B200_XOR, b[i] + bit, carry + i, t1 + i);
            }
            end_level();
            const bool last = (bit == nbits - 1);
            if (!last) {
                for (int i = 0; i < m; i++) gate(TFHE_B200_AND, t0 + i, t1 + i, t0 + i);  // main.cu:874
                end_level();
            }
            for (int i = 0; i < m; i++) {  // sum = a ^ t1, carry' = t0 ^ c  (main.cu:878)
                gate(TFHE_B200_XOR, a[i] + bit, t1 + i, out[i] + bit);
                if (!last) gate(TFHE_B200_XOR, t0 + i, carry + i, carry + i);
            }
            end_level();
        }
    }
};

int fail_msg(const char *m) {
    fprintf(stderr, "tfhe_b200 circuit: %s\n", m);
    return 1;
}

tfhe_b200_circuit *finish(tfhe_b200_circuit *c) {
    // upload the index tables and allocate the workspace
    const size_t ib = c->h_idx.size() * sizeof(int32_t);
    if (cudaMalloc(&c->d_idx, ib ? ib : 4) != cudaSuccess ||
        cudaMalloc(&c->d_ws, (size_t) c->nrows * c->words * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(c->d_idx, c->h_idx.data(), ib, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail_msg("device allocation failed");
        if (c->d_idx) cudaFree(c->d_idx);
        if (c->d_ws) cudaFree(c->d_ws);
        delete c;
        return nullptr;
    }
    return c;
}

}  // namespace

extern "C" {

int tfhe_b200_ctx_words(const tfhe_b200_ctx *ctx);  // engine.cu

// a + b for `count` pairs of nbits-bit integers (LSB first).  mode 0: bit-wise ripple carry,
// 1 + 3(nbits-1) - 1 levels; mode 1: number-wise carry-save iteration (taskLevelParallelAdd,
// main.cu:619-652): nbits levels of 2*nbits gates per number.
tfhe_b200_circuit *tfhe_b200_circuit_add(tfhe_b200_ctx *ctx, int nbits, int count, int mode) {
    if (!ctx || nbits < 1 || count < 1 || mode < 0 || mode > 1) return nullptr;
    tfhe_b200_circuit *c = new tfhe_b200_circuit();
    c->ctx = ctx;
    c->words = tfhe_b200_ctx_words(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_rows = count * nbits;
    if (mode == 0) {
        c->out_row0 = B.alloc(count * nbits);
        std::vector<int> va, vb, vo;
        for (int i = 0; i < count; i++) {
            va.push_back(a + i * nbits);
            vb.push_back(b + i * nbits);
            vo.push_back(c->out_row0 + i * nbits);
        }
        B.ripple_add(va, vb, vo, nbits);
    } else {
        c->row_zero = B.alloc(1);
        // ping-pong sets: (x = running sum, y = shifted carries)
        int x[2], y[2];
        for (int s = 0; s < 2; s++) {
            x[s] = B.alloc(count * nbits);
            y[s] = B.alloc(count * nbits);
        }
        int cur_x = a, cur_y = b;
        for (int round = 0; round < nbits; round++) {
            const int s = round & 1;
            for (int i = 0; i < count; i++)
                for (int bit = 0; bit < nbits; bit++) {
                    const int ra = cur_x + i * nbits + bit;
                    // the shifted-in low bit of the carry word is the constant 0
                    const int rb = (round > 0 && bit == 0) ? c->row_zero : cur_y + i * nbits + bit;
                    B.gate(TFHE_B200_XOR, ra, rb, x[s] + i * nbits + bit);
                    if (bit + 1 < nbits) B.gate(TFHE_B200_AND, ra, rb, y[s] + i * nbits + bit + 1);  // << 1
                }
            B.end_level();
            cur_x = x[s];
            cur_y = y[s];
        }
        c->out_row0 = cur_x;
    }
    return finish(c);
}

// a * b mod 2^nbits for `count` pairs (multiplyLweSamples main.cu:1483-1579, single precision;
// BOOTS_vectorMultiplication :1746): one AND level over the partial-product matrix, then a
// binary tree of lock-step ripple-carry adders.
static void build_mul(Builder &B, tfhe_b200_circuit *c, const std::vector<int> &a, const std::vector<int> &b,
                      const std::vector<int> &out, int nbits) {
    const int m = (int) a.size();
    // addend rows R[i][p]: (a << i) & b_i ; bits below i are the constant 0
    std::vector<std::vector<int>> R(nbits, std::vector<int>(m));
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++) R[i][p] = B.alloc(nbits);
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++)
            for (int k = 0; k < nbits; k++) {
                if (k >= i) B.gate(TFHE_B200_AND, a[p] + (k - i), b[p] + i, R[i][p] + k);  // main.cu:1524
            }
    B.end_level();
    // rows below the shift are copies of the zero constant: mark by pointing adders at row_zero.
    // (the workspace rows themselves are initialised to the constant at run time)
    int live = nbits;
    std::vector<std::vector<int>> cur = R;
    while (live > 1) {
        const int half = live / 2;
        std::vector<int> va, vb, vo;
        std::vector<std::vector<int>> next;
        for (int i = 0; i < half; i++) {
            std::vector<int> dst(m);
            for (int p = 0; p < m; p++) {
                dst[p] = (live == 2) ? out[p] : B.alloc(nbits);
                va.push_back(cur[i][p]);
                vb.push_back(cur[i + half][p]);
                vo.push_back(dst[p]);
            }
            next.push_back(dst);
        }
        B.ripple_add(va, vb, vo, nbits);  // main.cu:1549
        if (live & 1) next.push_back(cur[live - 1]);
        cur = next;
        live = (int) cur.size();
    }
    (void) c;
}

tfhe_b200_circuit *tfhe_b200_circuit_mul(tfhe_b200_ctx *ctx, int nbits, int count) {
    if (!ctx || nbits < 2 || count < 1) return nullptr;
    tfhe_b200_circuit *c = new tfhe_b200_circuit();
    c->ctx = ctx;
    c->words = tfhe_b200_ctx_words(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    c->row_zero = -2;  // whole workspace is initialised to the constant 0 at run time
    std::vector<int> va, vb, vo;
    for (int i = 0; i < count; i++) {
        va.push_back(a + i * nbits);
        vb.push_back(b + i * nbits);
        vo.push_back(c->out_row0 + i * nbits);
    }
    build_mul(B, c, va, vb, vo, nbits);
    return finish(c);
}

// C = A * B with A rows x inner, B inner x cols, elements nbits-bit integers mod 2^nbits
// (BOOTS_matrixMultiplication main.cu:2342-2462: all rows*cols*inner products as one vector
// multiplication, then a tree of vector additions over the inner index).
tfhe_b200_circuit *tfhe_b200_circuit_matmul(tfhe_b200_ctx *ctx, int rows, int inner, int cols, int nbits) {
    if (!ctx || rows < 1 || inner < 1 || cols < 1 || nbits < 2) return nullptr;
    tfhe_b200_circuit *c = new tfhe_b200_circuit();
    c->ctx = ctx;
    c->words = tfhe_b200_ctx_words(ctx);
    Builder B(c);
    const int A = B.operand(rows * inner * nbits), Bm = B.operand(inner * cols * nbits);
    c->out_row0 = B.alloc(rows * cols * nbits);
    c->out_rows = rows * cols * nbits;
    c->row_zero = -2;
    std::vector<int> va, vb, vo;
    std::vector<std::vector<int>> prod(inner);
    for (int k = 0; k < inner; k++)
        for (int r = 0; r < rows; r++)
            for (int q = 0; q < cols; q++) {
                va.push_back(A + (r * inner + k) * nbits);      // matMul_prepareLeftMat matrixUtility.cu:65
                vb.push_back(Bm + (k * cols + q) * nbits);      // matMul_prepareRightMat :82
                const int dst = (inner == 1) ? c->out_row0 + (r * cols + q) * nbits : B.alloc(nbits);
                vo.push_back(dst);
                prod[k].push_back(dst);
            }
    build_mul(B, c, va, vb, vo, nbits);
    int live = inner;
    while (live > 1) {
        const int half = live / 2;
        std::vector<int> xa, xb, xo;
        std::vector<std::vector<int>> next;
        for (int k = 0; k < half; k++) {
            std::vector<int> dst(rows * cols);
            for (int e = 0; e < rows * cols; e++) {
                dst[e] = (live == 2) ? c->out_row0 + e * nbits : B.alloc(nbits);
                xa.push_back(prod[k][e]);
                xb.push_back(prod[k + half][e]);
                xo.push_back(dst[e]);
            }
            next.push_back(dst);
        }
        B.ripple_add(xa, xb, xo, nbits);  // BOOTS_vectorAddition main.cu:1304
        if (live & 1) next.push_back(prod[live - 1]);
        prod = next;
        live = (int) prod.size();
    }
    return finish(c);
}

void tfhe_b200_circuit_destroy(tfhe_b200_circuit *c) {
    if (!c) return;
    if (c->d_idx) cudaFree(c->d_idx);
    if (c->d_ws) cudaFree(c->d_ws);
    delete c;
}

int tfhe_b200_circuit_levels(const tfhe_b200_circuit *c) { return c ? (int) c->levels.size() : 0; }
long long tfhe_b200_circuit_gates(const tfhe_b200_circuit *c) { return c ? c->n_gates : 0; }
int tfhe_b200_circuit_operands(const tfhe_b200_circuit *c) { return c ? (int) c->in_row0.size() : 0; }
int tfhe_b200_circuit_operand_rows(const tfhe_b200_circuit *c, int o) {
    return (c && o >= 0 && o < (int) c->in_rows.size()) ? c->in_rows[o] : 0;
}
int tfhe_b200_circuit_output_rows(const tfhe_b200_circuit *c) { return c ? c->out_rows : 0; }

// Runs the plan: operands[o] and d_out are DEVICE arrays of samples (rows of n+1 words).
int tfhe_b200_circuit_run(tfhe_b200_circuit *c, int32_t *d_out, const int32_t *const *operands, void *stream) {
    if (!c || !d_out || !operands) return fail_msg("null argument");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t rb = (size_t) c->words * sizeof(int32_t);
    if (c->row_zero == -2) {
        if (tfhe_b200_constant(c->ctx, c->d_ws, 0, c->nrows, stream)) return 1;
    } else if (c->row_zero >= 0) {
        if (tfhe_b200_constant(c->ctx, c->d_ws + (size_t) c->row_zero * c->words, 0, 1, stream)) return 1;
    }
    for (size_t o = 0; o < c->in_row0.size(); o++)
        if (cudaMemcpyAsync(c->d_ws + (size_t) c->in_row0[o] * c->words, operands[o], rb * c->in_rows[o],
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return fail_msg("operand copy failed");
    for (const Level &lv : c->levels) {
        tfhe_b200_gate_op ops[4];
        for (int i = 0; i < lv.nops; i++) {
            const Op &op = lv.ops[i];
            ops[i].gate = op.gate;
            ops[i].count = op.count;
            ops[i].a = ops[i].b = c->d_ws;
            ops[i].out = c->d_ws;
            ops[i].stride_a = ops[i].stride_b = ops[i].stride_out = c->words;
            ops[i].idx_a = c->d_idx + op.off_a;
            ops[i].idx_b = c->d_idx + op.off_b;
            ops[i].idx_out = c->d_idx + op.off_out;
        }
        if (tfhe_b200_gate_multi(c->ctx, ops, lv.nops, stream)) return 1;
    }
    if (cudaMemcpyAsync(d_out, c->d_ws + (size_t) c->out_row0 * c->words, rb * c->out_rows,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail_msg("result copy failed");
    return 0;
}

}  // extern "C"
