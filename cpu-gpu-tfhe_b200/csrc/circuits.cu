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
    int gate, count, off_a, off_b, off_out, off_c;
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
    std::vector<int> ga[TFHE_B200_NUM_GATES_EXT], gb[TFHE_B200_NUM_GATES_EXT], go[TFHE_B200_NUM_GATES_EXT],
        gc[TFHE_B200_NUM_GATES_EXT];

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
    void gate(int g, int a, int b, int out, int c3 = -1) {
        ga[g].push_back(a);
        gb[g].push_back(b);
        go[g].push_back(out);
        if (c3 >= 0) gc[g].push_back(c3);
    }

    // close the level: one bootstrap batch per (up to) 4 gate types
    void end_level() {
        Level lv;
        for (int g = 0; g < TFHE_B200_NUM_GATES_EXT; g++) {
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
            op.off_c = -1;
            if (!gc[g].empty()) {
                op.off_c = (int) c->h_idx.size();
                c->h_idx.insert(c->h_idx.end(), gc[g].begin(), gc[g].end());
            }
            gc[g].clear();
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

    // Parallel-prefix (Kogge-Stone) addition of m pairs: 2 + ceil(log2(nbits-1)) levels instead of
    // 3*nbits - 3 (the ripple schedules are bound by sequential depth).
    // The generate / propagate signals of a bit group are mutually exclusive, so the carry
    // operator G' = G | (P & G_prev) is ONE three-input threshold bootstrap (TFHE_B200_GPC).
    // Carry into bit i+1 = group generate of bits [0, i]; sum_i = p_i ^ carry_i.
    void prefix_add(const std::vector<int> &a, const std::vector<int> &b, const std::vector<int> &out, int nbits) {
        const int m = (int) a.size();
        const int np = nbits - 1;  // carry positions 0..nbits-2 (the carry out of the top bit is unused)
        const int p0 = alloc(m * nbits);
        const int g0 = alloc(m * (np > 0 ? np : 1));
        std::vector<int> G(m * (np > 0 ? np : 1)), P(m * (np > 0 ? np : 1));
        for (int i = 0; i < m; i++)
            for (int bit = 0; bit < nbits; bit++) {
                const int prow = (bit == 0) ? out[i] : p0 + i * nbits + bit;  // sum bit 0 is p_0 itself
                gate(TFHE_B200_XOR, a[i] + bit, b[i] + bit, prow);
                if (bit < np) {
                    gate(TFHE_B200_AND, a[i] + bit, b[i] + bit, g0 + i * np + bit);
                    G[i * np + bit] = g0 + i * np + bit;
                    P[i * np + bit] = prow;
                }
            }
        end_level();
        for (int d = 1; d < np; d *= 2) {
            // position `bit` holds the signals of bits [max(0, bit-d+1), bit]; after this level 2d bits
            std::vector<int> G2 = G, P2 = P;
            const bool more = 2 * d < np;
            for (int i = 0; i < m; i++)
                for (int bit = d; bit < np; bit++) {
                    const int r = i * np + bit;
                    G2[r] = alloc(1);
                    gate(TFHE_B200_GPC, G[r], P[r], G2[r], G[r - d]);
                    if (more && bit >= 2 * d) {  // groups that already reach bit 0 need no propagate
                        P2[r] = alloc(1);
                        gate(TFHE_B200_AND, P[r], P[r - d], P2[r]);
                    }
                }
            end_level();
            G.swap(G2);
            P.swap(P2);
        }
        if (nbits > 1) {
            for (int i = 0; i < m; i++)
                for (int bit = 1; bit < nbits; bit++)
                    gate(TFHE_B200_XOR, p0 + i * nbits + bit, G[i * np + bit - 1], out[i] + bit);
            end_level();
        }
    }

    void add(int adder, const std::vector<int> &a, const std::vector<int> &b, const std::vector<int> &out, int nbits) {
        if (adder == TFHE_B200_ADDER_PREFIX) prefix_add(a, b, out, nbits);
        else ripple_add(a, b, out, nbits);
    }
};

int fail_msg(const char *m) {
    fprintf(stderr, "tfhe_b200 circuit: %s\n", m);
    return 1;
}

// Plans are built on the host only; the index tables and the workspace go to the device at
// the first run (so that a plan can also be inspected / simulated without a GPU).
tfhe_b200_circuit *finish(tfhe_b200_circuit *c) { return c; }

int ensure_device(tfhe_b200_circuit *c) {
    if (c->d_idx != nullptr) return 0;
    const size_t ib = c->h_idx.size() * sizeof(int32_t);
    if (cudaMalloc(&c->d_idx, ib ? ib : 4) != cudaSuccess ||
        cudaMalloc(&c->d_ws, (size_t) c->nrows * c->words * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(c->d_idx, c->h_idx.data(), ib, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (c->d_idx) cudaFree(c->d_idx);
        if (c->d_ws) cudaFree(c->d_ws);
        c->d_idx = nullptr;
        c->d_ws = nullptr;
        return fail_msg("device allocation failed");
    }
    return 0;
}

// plaintext truth tables of the gate ids (boot-gates.cu:98-397 and the GPC extension)
int gate_truth(int g, int a, int b, int c3) {
    switch (g) {
        case TFHE_B200_NAND: return !(a && b);
        case TFHE_B200_OR: return a || b;
        case TFHE_B200_AND: return a && b;
        case TFHE_B200_XOR: return a ^ b;
        case TFHE_B200_XNOR: return !(a ^ b);
        case TFHE_B200_NOR: return !(a || b);
        case TFHE_B200_ANDNY: return !a && b;
        case TFHE_B200_ANDYN: return a && !b;
        case TFHE_B200_ORNY: return !a || b;
        case TFHE_B200_ORYN: return a || !b;
        case TFHE_B200_GPC: return a || (b && c3);
        default: return 0;
    }
}

tfhe_b200_circuit *new_plan(tfhe_b200_ctx *ctx) {
    tfhe_b200_circuit *c = new tfhe_b200_circuit();
    c->ctx = ctx;
    c->words = ctx ? tfhe_b200_ctx_words(ctx) : 0;
    return c;
}

}  // namespace

extern "C" {

// a + b for `count` pairs of nbits-bit integers (LSB first).  mode 0: bit-wise ripple carry,
// 3*nbits - 3 levels; mode 1: number-wise carry-save iteration (taskLevelParallelAdd,
// main.cu:619-652): nbits levels of 2*nbits gates per number; mode 2: parallel-prefix adder,
// 2 + ceil(log2(nbits-1)) levels (not in the reference).
tfhe_b200_circuit *tfhe_b200_circuit_add(tfhe_b200_ctx *ctx, int nbits, int count, int mode) {
    if (nbits < 1 || count < 1 || mode < 0 || mode > 2) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_rows = count * nbits;
    if (mode != 1) {
        c->out_row0 = B.alloc(count * nbits);
        std::vector<int> va, vb, vo;
        for (int i = 0; i < count; i++) {
            va.push_back(a + i * nbits);
            vb.push_back(b + i * nbits);
            vo.push_back(c->out_row0 + i * nbits);
        }
        B.add(mode == 2 ? TFHE_B200_ADDER_PREFIX : TFHE_B200_ADDER_RIPPLE, va, vb, vo, nbits);
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
// binary tree of lock-step adders (ripple carry as in the reference, or parallel prefix).
static void build_mul(Builder &B, const std::vector<int> &a, const std::vector<int> &b, const std::vector<int> &out,
                      int nbits, int adder) {
    const int m = (int) a.size();
    // addend rows R[i][p]: (a << i) & b_i ; bits below i are the constant 0 (the whole workspace
    // is initialised to the constant at run time, those rows are never written)
    std::vector<std::vector<int>> R(nbits, std::vector<int>(m));
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++) R[i][p] = B.alloc(nbits);
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++)
            for (int k = 0; k < nbits; k++) {
                if (k >= i) B.gate(TFHE_B200_AND, a[p] + (k - i), b[p] + i, R[i][p] + k);  // main.cu:1524
            }
    B.end_level();
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
        B.add(adder, va, vb, vo, nbits);  // main.cu:1549
        if (live & 1) next.push_back(cur[live - 1]);
        cur = next;
        live = (int) cur.size();
    }
}

tfhe_b200_circuit *tfhe_b200_circuit_mul_ex(tfhe_b200_ctx *ctx, int nbits, int count, int adder) {
    if (nbits < 2 || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
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
    build_mul(B, va, vb, vo, nbits, adder);
    return finish(c);
}

tfhe_b200_circuit *tfhe_b200_circuit_mul(tfhe_b200_ctx *ctx, int nbits, int count) {
    return tfhe_b200_circuit_mul_ex(ctx, nbits, count, TFHE_B200_ADDER_RIPPLE);
}

// C = A * B with A rows x inner, B inner x cols, elements nbits-bit integers mod 2^nbits
// (BOOTS_matrixMultiplication main.cu:2342-2462: all rows*cols*inner products as one vector
// multiplication, then a tree of vector additions over the inner index).
tfhe_b200_circuit *tfhe_b200_circuit_matmul_ex(tfhe_b200_ctx *ctx, int rows, int inner, int cols, int nbits,
                                               int adder) {
    if (rows < 1 || inner < 1 || cols < 1 || nbits < 2 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
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
    build_mul(B, va, vb, vo, nbits, adder);
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
        B.add(adder, xa, xb, xo, nbits);  // BOOTS_vectorAddition main.cu:1304
        if (live & 1) next.push_back(prod[live - 1]);
        prod = next;
        live = (int) prod.size();
    }
    return finish(c);
}

tfhe_b200_circuit *tfhe_b200_circuit_matmul(tfhe_b200_ctx *ctx, int rows, int inner, int cols, int nbits) {
    return tfhe_b200_circuit_matmul_ex(ctx, rows, inner, cols, nbits, TFHE_B200_ADDER_RIPPLE);
}

// Evaluates the plan on PLAINTEXT bits on the host (one int per sample row): the schedule's
// logic can be checked without keys or a GPU.  operand_bits[o]: operand_rows(o) ints in {0,1}.
int tfhe_b200_circuit_simulate(const tfhe_b200_circuit *c, int32_t *out_bits, const int32_t *const *operand_bits) {
    if (!c || !out_bits || !operand_bits) return fail_msg("null argument");
    std::vector<int32_t> ws((size_t) c->nrows, 0);
    for (size_t o = 0; o < c->in_row0.size(); o++)
        for (int r = 0; r < c->in_rows[o]; r++) ws[c->in_row0[o] + r] = operand_bits[o][r] & 1;
    for (const Level &lv : c->levels) {
        // all gates of a level read the state before the level (they run as one batch)
        std::vector<std::pair<int, int32_t>> writes;
        for (int i = 0; i < lv.nops; i++) {
            const Op &op = lv.ops[i];
            for (int g = 0; g < op.count; g++) {
                const int a = ws[c->h_idx[op.off_a + g]], b = ws[c->h_idx[op.off_b + g]];
                const int c3 = op.off_c >= 0 ? ws[c->h_idx[op.off_c + g]] : 0;
                writes.emplace_back(c->h_idx[op.off_out + g], gate_truth(op.gate, a, b, c3));
            }
        }
        for (const auto &w : writes) ws[w.first] = w.second;
    }
    for (int r = 0; r < c->out_rows; r++) out_bits[r] = ws[c->out_row0 + r];
    return 0;
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
    if (!c->ctx) return fail_msg("plan was built without an engine context (simulation only)");
    if (ensure_device(c)) return 1;
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
            ops[i].c = op.off_c >= 0 ? c->d_ws : nullptr;
            ops[i].stride_c = c->words;
            ops[i].idx_c = op.off_c >= 0 ? c->d_idx + op.off_c : nullptr;
        }
        if (tfhe_b200_gate_multi(c->ctx, ops, lv.nops, stream)) return 1;
    }
    if (cudaMemcpyAsync(d_out, c->d_ws + (size_t) c->out_row0 * c->words, rb * c->out_rows,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail_msg("result copy failed");
    return 0;
}

}  // extern "C"
