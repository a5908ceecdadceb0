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

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "kernels.h"

namespace {

// One launch group of a plan.
//   GATES : up to 4 runs of two-/three-input bootstrapped gates (one blind-rotate + one key-switch launch)
//   MUX   : bootsMUX gates (two bootstraps + one key switch each, boot-gates.cu:407-448), one launch pair
//   LINEAR: bootstrap-free row operations (bootsCOPY / bootsNOT / bootsCONSTANT, boot-gates.cu:242-267)
enum { LV_GATES = 0, LV_MUX = 1, LV_LINEAR = 2 };
enum { LIN_COPY = 0, LIN_NOT = 1, LIN_ZERO = 2, LIN_ONE = 3, LIN_KINDS = 4 };

struct Op {
    int gate, count, off_a, off_b, off_out, off_c;
    int off_d = -1;  // fourth operand (TFHE_B200_SUMC)
};

struct Level {
    int type = LV_GATES;
    Op ops[4];   // GATES: runs; MUX: ops[0] (a = selector, b, c = off_c); LINEAR: ops[0] (gate = LIN_*, a = source)
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
    int depth = 0;                      // sequential bootstrap batches (GATES and MUX groups)
    std::vector<int32_t> h_idx;
    int32_t *d_idx = nullptr;
    int32_t *d_ws = nullptr;
    int32_t *d_u = nullptr;             // extracted samples of the widest level (plan-owned scratch)
    size_t u_bytes = 0;
    long long n_gates = 0;
    // CUDA graph of the launch sequence (all levels; operand / result copies stay outside, so the graph
    // does not depend on the caller's buffers)
    bool graph_enabled = true, graph_failed = false, used_graph = false;
    cudaGraphExec_t graph_exec = nullptr;
    unsigned long long graph_launches = 0;  // kernel launches one replay stands for
};

namespace {

typedef std::vector<int> Bits;  // rows of the bits of one integer, LSB first

// ---- plan builder -----------------------------------------------------------------------
// Ops are collected per LEVEL.  Inside a level the groups run in the order GATES, MUX, LINEAR;
// gates of one group read the state before the group (a group may overwrite its own operands),
// later groups of the same level see the results of earlier ones.

struct Builder {
    tfhe_b200_circuit *c;
    // gates of the level under construction, grouped by gate type
    std::vector<int> ga[TFHE_B200_NUM_GATES_EXT], gb[TFHE_B200_NUM_GATES_EXT], go[TFHE_B200_NUM_GATES_EXT],
        gc[TFHE_B200_NUM_GATES_EXT], gd[TFHE_B200_NUM_GATES_EXT];
    std::vector<int> ms, mb, mc, mo;                  // MUX: selector, then-value, else-value, out
    std::vector<int> li[LIN_KINDS], lo[LIN_KINDS];    // LINEAR

    explicit Builder(tfhe_b200_circuit *circ) : c(circ) {}

    int alloc(int rows = 1) {
        const int r = c->nrows;
        c->nrows += rows;
        return r;
    }

    Bits alloc_bits(int n) {
        Bits v(n);
        const int r = alloc(n);
        for (int i = 0; i < n; i++) v[i] = r + i;
        return v;
    }

    static Bits bits_at(int row0, int n) {
        Bits v(n);
        for (int i = 0; i < n; i++) v[i] = row0 + i;
        return v;
    }

    int operand(int rows) {
        const int r = alloc(rows);
        c->in_row0.push_back(r);
        c->in_rows.push_back(rows);
        return r;
    }

    // schedule out = gate(a, b [, c3]) in the current level
    void gate(int g, int a, int b, int out, int c3 = -1, int c4 = -1) {
        ga[g].push_back(a);
        gb[g].push_back(b);
        go[g].push_back(out);
        if (c3 >= 0) gc[g].push_back(c3);
        if (c4 >= 0) gd[g].push_back(c4);
    }
    void mux(int sel, int b, int cc, int out) {  // out = sel ? b : cc
        ms.push_back(sel);
        mb.push_back(b);
        mc.push_back(cc);
        mo.push_back(out);
    }
    void lin(int kind, int in, int out) {
        li[kind].push_back(in);
        lo[kind].push_back(out);
    }
    void copy(int in, int out) { lin(LIN_COPY, in, out); }
    void not_(int in, int out) { lin(LIN_NOT, in, out); }
    void constant(int value, int out) { lin(value ? LIN_ONE : LIN_ZERO, 0, out); }

    int push_idx(const std::vector<int> &v) {
        const int off = (int) c->h_idx.size();
        c->h_idx.insert(c->h_idx.end(), v.begin(), v.end());
        return off;
    }

    // close the level
    void end_level() {
        Level lv;
        for (int g = 0; g < TFHE_B200_NUM_GATES_EXT; g++) {
            if (ga[g].empty()) continue;
            if (lv.nops == 4) {
                c->levels.push_back(lv);
                c->depth++;
                lv = Level();
            }
            Op &op = lv.ops[lv.nops++];
            op.gate = g;
            op.count = (int) ga[g].size();
            op.off_a = push_idx(ga[g]);
            op.off_b = push_idx(gb[g]);
            op.off_out = push_idx(go[g]);
            op.off_c = gc[g].empty() ? -1 : push_idx(gc[g]);
            op.off_d = gd[g].empty() ? -1 : push_idx(gd[g]);
            c->n_gates += op.count;
            ga[g].clear();
            gb[g].clear();
            go[g].clear();
            gc[g].clear();
            gd[g].clear();
        }
        if (lv.nops) {
            c->levels.push_back(lv);
            c->depth++;
        }
        if (!ms.empty()) {
            Level m;
            m.type = LV_MUX;
            m.nops = 1;
            m.ops[0].gate = -1;
            m.ops[0].count = (int) ms.size();
            m.ops[0].off_a = push_idx(ms);
            m.ops[0].off_b = push_idx(mb);
            m.ops[0].off_c = push_idx(mc);
            m.ops[0].off_out = push_idx(mo);
            c->n_gates += m.ops[0].count;
            c->levels.push_back(m);
            c->depth++;
            ms.clear();
            mb.clear();
            mc.clear();
            mo.clear();
        }
        for (int k = 0; k < LIN_KINDS; k++) {
            if (lo[k].empty()) continue;
            Level l;
            l.type = LV_LINEAR;
            l.nops = 1;
            l.ops[0].gate = k;
            l.ops[0].count = (int) lo[k].size();
            l.ops[0].off_a = push_idx(li[k]);
            l.ops[0].off_b = -1;
            l.ops[0].off_c = -1;
            l.ops[0].off_out = push_idx(lo[k]);
            c->levels.push_back(l);
            li[k].clear();
            lo[k].clear();
        }
    }

    // ---- adders ---------------------------------------------------------------------------
    // All adders work on m independent pairs in lock-step; a[i], b[i], out[i] are the bit rows of
    // pair i.  cin1: carry-in 1 (used for a - b = a + ~b + 1).  Result truncated to nbits.

    // Ripple-carry addition (taskLevelParallelAdd_bitwise[_vector_coalInput], main.cu:821-890 /
    // 1138-1302; the 5-gate full adder of Cipher::addBits, Cipher.cu:367-378, as 3 levels of
    // 2/1/2 gates).
    void ripple_add(const std::vector<Bits> &a, const std::vector<Bits> &b, const std::vector<Bits> &out, int nbits,
                    bool cin1 = false) {
        const int m = (int) a.size();
        const int carry = alloc(m), t0 = alloc(m), t1 = alloc(m);
        for (int i = 0; i < m; i++) {  // bit 0: (carry, sum) = (AND, XOR), bootsANDXOR main.cu:849
            if (nbits > 1) gate(cin1 ? TFHE_B200_OR : TFHE_B200_AND, a[i][0], b[i][0], carry + i);
            gate(cin1 ? TFHE_B200_XNOR : TFHE_B200_XOR, a[i][0], b[i][0], out[i][0]);
        }
        end_level();
        for (int bit = 1; bit < nbits; bit++) {
            for (int i = 0; i < m; i++) {  // t0 = a ^ c, t1 = b ^ c  (bootsXORXOR main.cu:869)
                gate(TFHE_B200_XOR, a[i][bit], carry + i, t0 + i);
                gate(TFHE_B200_XOR, b[i][bit], carry + i, t1 + i);
            }
            end_level();
            const bool last = (bit == nbits - 1);
            if (!last) {
                for (int i = 0; i < m; i++) gate(TFHE_B200_AND, t0 + i, t1 + i, t0 + i);  // main.cu:874
                end_level();
            }
            for (int i = 0; i < m; i++) {  // sum = a ^ t1, carry' = t0 ^ c  (main.cu:878)
                gate(TFHE_B200_XOR, a[i][bit], t1 + i, out[i][bit]);
                if (!last) gate(TFHE_B200_XOR, t0 + i, carry + i, carry + i);
            }
            end_level();
        }
    }

    // Kogge-Stone scan of (generate, propagate) pairs, LSB first: on return G[i][k] is the
    // generate signal of bits [0, k] (the carry into bit k+1).  The generate / propagate signals
    // of a bit group are mutually exclusive, so the carry operator G' = G | (P & G_prev) is ONE
    // three-input threshold bootstrap (TFHE_B200_GPC).  np positions per number.
    // fuse_prop / fuse_out (optional): the LAST level does not produce the carries but, fused with the
    // carry operator, the sum bits themselves: out[k+1] = prop[k+1] ^ carry_k in ONE bootstrap
    // (TFHE_B200_SUMC: prop ^ (G | (P & G_prev))), or prop[k+1] ^ G[k] for the carries that are final
    // already — so an addition ends with its last scan level instead of one level after it.
    void prefix_scan(std::vector<Bits> &G, std::vector<Bits> &P, int np, const std::vector<Bits> *fuse_prop = nullptr,
                     const std::vector<Bits> *fuse_out = nullptr) {
        const int m = (int) G.size();
        for (int d = 1; d < np; d *= 2) {
            if (fuse_prop != nullptr && 2 * d >= np) {  // last level
                for (int i = 0; i < m; i++)
                    for (int k = 0; k < np; k++) {
                        const int prow = (*fuse_prop)[i][k + 1], orow = (*fuse_out)[i][k + 1];
                        if (k >= d) gate(TFHE_B200_SUMC, prow, G[i][k], orow, P[i][k], G[i][k - d]);
                        else gate(TFHE_B200_XOR, prow, G[i][k], orow);
                    }
                end_level();
                return;
            }
            // position k holds the signals of bits [max(0, k-d+1), k]; after this level 2d bits
            std::vector<Bits> G2 = G, P2 = P;
            const bool more = 2 * d < np;
            for (int i = 0; i < m; i++)
                for (int k = d; k < np; k++) {
                    G2[i][k] = alloc(1);
                    gate(TFHE_B200_GPC, G[i][k], P[i][k], G2[i][k], G[i][k - d]);
                    if (more && k >= 2 * d) {  // groups that already reach bit 0 need no propagate
                        P2[i][k] = alloc(1);
                        gate(TFHE_B200_AND, P[i][k], P[i][k - d], P2[i][k]);
                    }
                }
            end_level();
            G.swap(G2);
            P.swap(P2);
        }
    }

    // Parallel-prefix addition: 1 + ceil(log2(nbits-1)) levels instead of 3*nbits - 3 (the last scan level
    // produces the sum bits directly, TFHE_B200_SUMC)
    // (the ripple schedules are bound by sequential depth).
    void prefix_add(const std::vector<Bits> &a, const std::vector<Bits> &b, const std::vector<Bits> &out, int nbits,
                    bool cin1 = false) {
        const int m = (int) a.size();
        const int np = nbits - 1;  // carries into bits 1..nbits-1 (the carry out of the top bit is unused)
        std::vector<Bits> G(m), P(m), prop(m);
        for (int i = 0; i < m; i++) {
            prop[i] = alloc_bits(nbits);
            G[i] = alloc_bits(np > 0 ? np : 1);
            P[i].assign(np > 0 ? np : 1, -1);
            for (int bit = 0; bit < nbits; bit++) {
                // sum bit 0 is p_0 itself (carry-in 0) or its complement (carry-in 1)
                const int prow = (bit == 0) ? out[i][0] : prop[i][bit];
                gate((bit == 0 && cin1) ? TFHE_B200_XNOR : TFHE_B200_XOR, a[i][bit], b[i][bit], prow);
                if (bit < np) {
                    // with carry-in 1 bit 0 generates whenever it generates or propagates: a | b
                    gate((bit == 0 && cin1) ? TFHE_B200_OR : TFHE_B200_AND, a[i][bit], b[i][bit], G[i][bit]);
                    P[i][bit] = prow;  // P[i][0] is never read
                }
            }
        }
        end_level();
        if (np > 1) {  // at least one scan level: the last one produces the sum bits directly
            prefix_scan(G, P, np, &prop, &out);
        } else if (nbits > 1) {
            for (int i = 0; i < m; i++)
                for (int bit = 1; bit < nbits; bit++) gate(TFHE_B200_XOR, prop[i][bit], G[i][bit - 1], out[i][bit]);
            end_level();
        }
    }

    void add(int adder, const std::vector<Bits> &a, const std::vector<Bits> &b, const std::vector<Bits> &out,
             int nbits, bool cin1 = false) {
        if (adder == TFHE_B200_ADDER_PREFIX) prefix_add(a, b, out, nbits, cin1);
        else ripple_add(a, b, out, nbits, cin1);
    }

    // out = a - b = a + ~b + 1 (the reference: a + twosComplement(b), Cipher.cu:329-332)
    void sub(int adder, const std::vector<Bits> &a, const std::vector<Bits> &b, const std::vector<Bits> &out,
             int nbits) {
        const int m = (int) a.size();
        std::vector<Bits> nb(m);
        for (int i = 0; i < m; i++) {
            nb[i] = alloc_bits(nbits);
            for (int bit = 0; bit < nbits; bit++) not_(b[i][bit], nb[i][bit]);
        }
        end_level();
        add(adder, a, nb, out, nbits, true);
    }

    // Binary tree reduction of per-bit signals with a two-input gate; returns the row of the result.
    std::vector<int> reduce_tree(int g, std::vector<Bits> cur) {
        const int m = (int) cur.size();
        while (cur[0].size() > 1) {
            std::vector<Bits> next(m);
            for (int i = 0; i < m; i++) {
                const int n = (int) cur[i].size();
                for (int k = 0; k + 1 < n; k += 2) {
                    const int r = alloc(1);
                    gate(g, cur[i][k], cur[i][k + 1], r);
                    next[i].push_back(r);
                }
                if (n & 1) next[i].push_back(cur[i][n - 1]);
            }
            end_level();
            cur.swap(next);
        }
        std::vector<int> res(m);
        for (int i = 0; i < m; i++) res[i] = cur[i][0];
        return res;
    }

    // a > b (row per pair).  a > b  <=>  a + ~b carries out; the carry is the group generate of all
    // bits with g = a & ~b, p = a XNOR b, reduced in a tree with the same three-input carry operator.
    // Signed operands: comparing with the sign bits flipped, i.e. g = ~a & b at the top bit
    // (the reference chains compareBit_g over the bits and XORs the sign difference in,
    // Cipher.cu:574-584; same truth table, depth nbits instead of 1 + log2).
    std::vector<int> greater(const std::vector<Bits> &a, const std::vector<Bits> &b, int nbits, bool is_signed) {
        const int m = (int) a.size();
        std::vector<Bits> G(m), P(m);
        for (int i = 0; i < m; i++) {
            G[i] = alloc_bits(nbits);
            P[i] = alloc_bits(nbits);
            for (int bit = 0; bit < nbits; bit++) {
                const bool top = is_signed && bit == nbits - 1;
                gate(top ? TFHE_B200_ANDNY : TFHE_B200_ANDYN, a[i][bit], b[i][bit], G[i][bit]);
                if (bit > 0) gate(TFHE_B200_XNOR, a[i][bit], b[i][bit], P[i][bit]);  // P of bit 0 is never used
            }
        }
        end_level();
        // (G, P)[k] covers a group of bits; neighbours merge as hi o lo = (G_hi | P_hi & G_lo, P_hi & P_lo)
        while (G[0].size() > 1) {
            std::vector<Bits> G2(m), P2(m);
            for (int i = 0; i < m; i++) {
                const int n = (int) G[i].size();
                for (int k = 0; k + 1 < n; k += 2) {
                    const int g = alloc(1);
                    gate(TFHE_B200_GPC, G[i][k + 1], P[i][k + 1], g, G[i][k]);
                    G2[i].push_back(g);
                    if (k > 0) {  // the group containing bit 0 never needs its propagate
                        const int pr = alloc(1);
                        gate(TFHE_B200_AND, P[i][k + 1], P[i][k], pr);
                        P2[i].push_back(pr);
                    } else {
                        P2[i].push_back(-1);
                    }
                }
                if (n & 1) {
                    G2[i].push_back(G[i][n - 1]);
                    P2[i].push_back(P[i][n - 1]);
                }
            }
            end_level();
            G.swap(G2);
            P.swap(P2);
        }
        std::vector<int> res(m);
        for (int i = 0; i < m; i++) res[i] = G[i][0];
        return res;
    }

    // a == b: AND over the bits of a XNOR b (the reference ORs the XORs serially, Cipher.cu:600-616)
    std::vector<int> equal(const std::vector<Bits> &a, const std::vector<Bits> &b, int nbits) {
        const int m = (int) a.size();
        std::vector<Bits> e(m);
        for (int i = 0; i < m; i++) {
            e[i] = alloc_bits(nbits);
            for (int bit = 0; bit < nbits; bit++) gate(TFHE_B200_XNOR, a[i][bit], b[i][bit], e[i][bit]);
        }
        end_level();
        return reduce_tree(TFHE_B200_AND, e);
    }

    // Carry-save reduction.  cols[i][k] = rows of the bits of weight 2^k that add up to number i
    // (k < nbits; carries out of the top column are dropped: arithmetic mod 2^nbits).  Every level
    // compresses each column's bits in triples with full adders — sum bit stays, carry bit moves one
    // column up; both are ONE bootstrap each (TFHE_B200_XOR3 / TFHE_B200_MAJ) and all full adders of
    // a level form one batch — until no column holds more than two bits; the two remaining rows go
    // through one addition (`adder`).  zero_row: an encryption of 0 that is never written.
    void carry_save_sum(std::vector<std::vector<std::vector<int>>> cols, const std::vector<Bits> &out, int nbits,
                        int adder, int zero_row) {
        const int m = (int) cols.size();
        // Dadda's height sequence 2, 3, 4, 6, 9, 13, 19, 28, ... (d' = floor(3d/2)): every stage brings all
        // columns down to the next target with as few adders as possible, so the number of stages is
        // logarithmic in the tallest column even for ragged (triangular) bit matrices, where compressing
        // greedily lets a carry ripple one column per level.
        size_t tallest = 0;
        for (int i = 0; i < m; i++)
            for (int k = 0; k < nbits; k++) tallest = std::max(tallest, cols[i][k].size());
        std::vector<size_t> targets;
        for (size_t d = 2; d < tallest; d = d * 3 / 2) targets.push_back(d);
        for (size_t st = targets.size(); st-- > 0;) {
            const size_t d = targets[st];
            std::vector<std::vector<std::vector<int>>> next(m, std::vector<std::vector<int>>(nbits));
            for (int i = 0; i < m; i++) {
                size_t carries_in = 0;  // carries produced by column k-1 in this stage
                for (int k = 0; k < nbits; k++) {
                    const std::vector<int> &col = cols[i][k];
                    size_t used = 0, height = col.size() + carries_in, carries_out = 0;
                    while (height > d) {
                        if (height == d + 1 && used + 2 <= col.size()) {  // half adder: XOR + AND
                            const int sum = alloc(1);
                            gate(TFHE_B200_XOR, col[used], col[used + 1], sum);
                            next[i][k].push_back(sum);
                            if (k + 1 < nbits) {
                                const int carry = alloc(1);
                                gate(TFHE_B200_AND, col[used], col[used + 1], carry);
                                next[i][k + 1].push_back(carry);
                            }
                            used += 2;
                            height -= 1;
                        } else {  // full adder: XOR3 + MAJ
                            const int sum = alloc(1);
                            gate(TFHE_B200_XOR3, col[used], col[used + 1], sum, col[used + 2]);
                            next[i][k].push_back(sum);
                            if (k + 1 < nbits) {
                                const int carry = alloc(1);
                                gate(TFHE_B200_MAJ, col[used], col[used + 1], carry, col[used + 2]);
                                next[i][k + 1].push_back(carry);
                            }
                            used += 3;
                            height -= 2;
                        }
                        carries_out++;
                    }
                    for (; used < col.size(); used++) next[i][k].push_back(col[used]);
                    carries_in = carries_out;
                }
            }
            end_level();
            cols.swap(next);
        }
        // two rows left: x + y (columns with one or no bit are padded with the constant 0)
        std::vector<Bits> x(m, Bits(nbits, zero_row)), y(m, Bits(nbits, zero_row));
        bool need_add = false;
        for (int i = 0; i < m; i++)
            for (int k = 0; k < nbits; k++) {
                if (cols[i][k].size() >= 1) x[i][k] = cols[i][k][0];
                if (cols[i][k].size() >= 2) {
                    y[i][k] = cols[i][k][1];
                    need_add = true;
                }
            }
        if (need_add) {
            add(adder == TFHE_B200_ADDER_RIPPLE ? TFHE_B200_ADDER_RIPPLE : TFHE_B200_ADDER_PREFIX, x, y, out, nbits);
        } else {
            for (int i = 0; i < m; i++)
                for (int k = 0; k < nbits; k++) copy(x[i][k], out[i][k]);
            end_level();
        }
    }

    // out = -a (twosComplement, Cipher.cu:286-298: out_i = a_i ^ (a_0 | ... | a_{i-1})), with the
    // running OR computed as a Kogge-Stone scan instead of a serial chain.
    void negate(const std::vector<Bits> &a, const std::vector<Bits> &out, int nbits) {
        const int m = (int) a.size();
        std::vector<Bits> r = a;  // r[k] = OR of bits [max(0, k-d+1), k]
        for (int d = 1; d < nbits - 1; d *= 2) {
            std::vector<Bits> r2 = r;
            for (int i = 0; i < m; i++)
                for (int k = d; k < nbits - 1; k++) {
                    r2[i][k] = alloc(1);
                    gate(TFHE_B200_OR, r[i][k], r[i][k - d], r2[i][k]);
                }
            end_level();
            r.swap(r2);
        }
        for (int i = 0; i < m; i++) {
            copy(a[i][0], out[i][0]);
            for (int bit = 1; bit < nbits; bit++) gate(TFHE_B200_XOR, a[i][bit], r[i][bit - 1], out[i][bit]);
        }
        end_level();
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
    // widest bootstrap batch of the plan: its extracted samples (N+1 words each) need scratch
    long long widest = 1;
    for (const Level &lv : c->levels) {
        long long w = 0;
        for (int i = 0; i < lv.nops; i++) w += lv.ops[i].count;
        if (lv.type == LV_MUX) w *= 2;
        if (lv.type != LV_LINEAR && w > widest) widest = w;
    }
    c->u_bytes = (size_t) widest * (tfhe_b200::kKsRowWords * 2 + 1) * sizeof(int32_t);  // N + 1 = 1025 words
    if (cudaSetDevice(tfhe_b200_ctx_device(c->ctx)) != cudaSuccess || cudaMalloc(&c->d_idx, ib ? ib : 4) != cudaSuccess ||
        cudaMalloc(&c->d_ws, (size_t) c->nrows * c->words * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&c->d_u, c->u_bytes) != cudaSuccess ||
        cudaMemcpy(c->d_idx, c->h_idx.data(), ib, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (c->d_idx) cudaFree(c->d_idx);
        if (c->d_ws) cudaFree(c->d_ws);
        if (c->d_u) cudaFree(c->d_u);
        c->d_idx = nullptr;
        c->d_ws = nullptr;
        c->d_u = nullptr;
        return fail_msg("device allocation failed");
    }
    return 0;
}

// plaintext truth tables of the gate ids (boot-gates.cu:98-397 and the GPC extension)
int gate_truth(int g, int a, int b, int c3, int c4 = 0) {
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
        case TFHE_B200_XOR3: return a ^ b ^ c3;
        case TFHE_B200_MAJ: return (a + b + c3) >= 2;
        case TFHE_B200_SUMC: return a ^ (b || (c3 && c4));
        default: return 0;
    }
}

tfhe_b200_circuit *new_plan(tfhe_b200_ctx *ctx) {
    tfhe_b200_circuit *c = new tfhe_b200_circuit();
    c->ctx = ctx;
    c->words = ctx ? tfhe_b200_ctx_words(ctx) : 0;
    return c;
}

// operand / result rows of `count` nbits-bit integers stored contiguously from row0
std::vector<Bits> numbers_at(int row0, int count, int nbits) {
    std::vector<Bits> v(count);
    for (int i = 0; i < count; i++) v[i] = Builder::bits_at(row0 + i * nbits, nbits);
    return v;
}

}  // namespace

extern "C" {

// a + b for `count` pairs of nbits-bit integers (LSB first).  mode 0: bit-wise ripple carry,
// 3*nbits - 3 levels; mode 1: number-wise carry-save iteration (taskLevelParallelAdd,
// main.cu:619-652): nbits levels of 2*nbits gates per number; mode 2: parallel-prefix adder,
// 1 + ceil(log2(nbits-1)) levels (not in the reference).
tfhe_b200_circuit *tfhe_b200_circuit_add(tfhe_b200_ctx *ctx, int nbits, int count, int mode) {
    if (nbits < 1 || count < 1 || mode < 0 || mode > 2) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_rows = count * nbits;
    if (mode != 1) {
        c->out_row0 = B.alloc(count * nbits);
        B.add(mode == 2 ? TFHE_B200_ADDER_PREFIX : TFHE_B200_ADDER_RIPPLE, numbers_at(a, count, nbits),
              numbers_at(b, count, nbits), numbers_at(c->out_row0, count, nbits), nbits);
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
// partial-product bits of a * b mod 2^nbits appended to the columns of one number: ONE AND gate each
static void partial_products(Builder &B, const Bits &a, const Bits &b, int nbits,
                             std::vector<std::vector<int>> &cols) {
    for (int i = 0; i < nbits; i++)
        for (int k = i; k < nbits; k++) {
            const int r = B.alloc(1);
            B.gate(TFHE_B200_AND, a[k - i], b[i], r);
            cols[k].push_back(r);
        }
}

static void build_mul(Builder &B, const std::vector<Bits> &a, const std::vector<Bits> &b,
                      const std::vector<Bits> &out, int nbits, int adder) {
    const int m = (int) a.size();
    if (adder == TFHE_B200_ADDER_CARRY_SAVE) {
        const int zero_row = B.alloc(1);  // never written: the workspace is initialised to the constant 0
        std::vector<std::vector<std::vector<int>>> cols(m, std::vector<std::vector<int>>(nbits));
        for (int p = 0; p < m; p++) partial_products(B, a[p], b[p], nbits, cols[p]);
        B.end_level();
        B.carry_save_sum(cols, out, nbits, TFHE_B200_ADDER_PREFIX, zero_row);
        return;
    }
    // addend rows R[i][p]: (a << i) & b_i ; bits below i are the constant 0 (the whole workspace
    // is initialised to the constant at run time, those rows are never written)
    std::vector<std::vector<Bits>> R(nbits, std::vector<Bits>(m));
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++) R[i][p] = B.alloc_bits(nbits);
    for (int i = 0; i < nbits; i++)
        for (int p = 0; p < m; p++)
            for (int k = i; k < nbits; k++) B.gate(TFHE_B200_AND, a[p][k - i], b[p][i], R[i][p][k]);  // main.cu:1524
    B.end_level();
    int live = nbits;
    std::vector<std::vector<Bits>> cur = R;
    while (live > 1) {
        const int half = live / 2;
        std::vector<Bits> va, vb, vo;
        std::vector<std::vector<Bits>> next;
        for (int i = 0; i < half; i++) {
            std::vector<Bits> dst(m);
            for (int p = 0; p < m; p++) {
                dst[p] = (live == 2) ? out[p] : B.alloc_bits(nbits);
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
    if (nbits < 2 || count < 1 || adder < 0 || adder > 2) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    c->row_zero = -2;  // whole workspace is initialised to the constant 0 at run time
    build_mul(B, numbers_at(a, count, nbits), numbers_at(b, count, nbits), numbers_at(c->out_row0, count, nbits), nbits,
              adder);
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
    if (rows < 1 || inner < 1 || cols < 1 || nbits < 2 || adder < 0 || adder > 2) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int A = B.operand(rows * inner * nbits), Bm = B.operand(inner * cols * nbits);
    c->out_row0 = B.alloc(rows * cols * nbits);
    c->out_rows = rows * cols * nbits;
    c->row_zero = -2;
    if (adder == TFHE_B200_ADDER_CARRY_SAVE) {
        // every element of C is ONE carry-save sum over the partial-product bits of all its `inner` products
        const int zero_row = B.alloc(1);
        std::vector<std::vector<std::vector<int>>> colsv(rows * cols, std::vector<std::vector<int>>(nbits));
        for (int r = 0; r < rows; r++)
            for (int q = 0; q < cols; q++)
                for (int k = 0; k < inner; k++)
                    partial_products(B, Builder::bits_at(A + (r * inner + k) * nbits, nbits),
                                     Builder::bits_at(Bm + (k * cols + q) * nbits, nbits), nbits, colsv[r * cols + q]);
        B.end_level();
        B.carry_save_sum(colsv, numbers_at(c->out_row0, rows * cols, nbits), nbits, TFHE_B200_ADDER_PREFIX, zero_row);
        return finish(c);
    }
    std::vector<Bits> va, vb, vo;
    std::vector<std::vector<Bits>> prod(inner);
    for (int k = 0; k < inner; k++)
        for (int r = 0; r < rows; r++)
            for (int q = 0; q < cols; q++) {
                va.push_back(Builder::bits_at(A + (r * inner + k) * nbits, nbits));   // matMul_prepareLeftMat matrixUtility.cu:65
                vb.push_back(Builder::bits_at(Bm + (k * cols + q) * nbits, nbits));   // matMul_prepareRightMat :82
                const Bits dst = (inner == 1) ? Builder::bits_at(c->out_row0 + (r * cols + q) * nbits, nbits)
                                              : B.alloc_bits(nbits);
                vo.push_back(dst);
                prod[k].push_back(dst);
            }
    build_mul(B, va, vb, vo, nbits, adder);
    int live = inner;
    while (live > 1) {
        const int half = live / 2;
        std::vector<Bits> xa, xb, xo;
        std::vector<std::vector<Bits>> next;
        for (int k = 0; k < half; k++) {
            std::vector<Bits> dst(rows * cols);
            for (int e = 0; e < rows * cols; e++) {
                dst[e] = (live == 2) ? Builder::bits_at(c->out_row0 + e * nbits, nbits) : B.alloc_bits(nbits);
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

// Full (double-precision) product of w-bit operands into 2w rows
// (multiplyLweSamples / BOOTS_vectorMultiplication with isDoublePrecision, main.cu:1483, 1746):
// addend i = (a & b_i) << i as a 2w-bit number whose other bits ARE the constant-zero row.
static void build_mul_full(Builder &B, const std::vector<Bits> &a, const std::vector<Bits> &b,
                           const std::vector<Bits> &out, int w, int adder, int zero_row) {
    const int m = (int) a.size(), W = 2 * w;
    std::vector<std::vector<Bits>> cur(w, std::vector<Bits>(m));
    for (int i = 0; i < w; i++)
        for (int p = 0; p < m; p++) {
            cur[i][p].assign(W, zero_row);
            for (int k = 0; k < w; k++) {
                const int r = B.alloc(1);
                B.gate(TFHE_B200_AND, a[p][k], b[p][i], r);
                cur[i][p][i + k] = r;
            }
        }
    B.end_level();
    int live = w;
    while (live > 1) {
        const int half = live / 2;
        std::vector<Bits> va, vb, vo;
        std::vector<std::vector<Bits>> next;
        for (int i = 0; i < half; i++) {
            std::vector<Bits> dst(m);
            for (int p = 0; p < m; p++) {
                dst[p] = (live == 2) ? out[p] : B.alloc_bits(W);
                va.push_back(cur[i][p]);
                vb.push_back(cur[i + half][p]);
                vo.push_back(dst[p]);
            }
            next.push_back(dst);
        }
        B.add(adder, va, vb, vo, W);
        if (live & 1) next.push_back(cur[live - 1]);
        cur = next;
        live = (int) cur.size();
    }
    if (w == 1)
        for (int p = 0; p < m; p++) {
            B.copy(cur[0][p][0], out[p][0]);
            B.constant(0, out[p][1]);
        }
    if (w == 1) B.end_level();
}

tfhe_b200_circuit *tfhe_b200_circuit_mul_full(tfhe_b200_ctx *ctx, int nbits, int count, int adder) {
    if (nbits < 1 || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * 2 * nbits);
    c->out_rows = count * 2 * nbits;
    c->row_zero = B.alloc(1);
    build_mul_full(B, numbers_at(a, count, nbits), numbers_at(b, count, nbits), numbers_at(c->out_row0, count, 2 * nbits),
                   nbits, adder, c->row_zero);
    return finish(c);
}

// One level of Karatsuba (karatMasterSuba, main.cu:1866-2087): with h = nbits/2,
//   X = Xl + 2^h Xr, Y = Yl + 2^h Yr,  P1 = Xl*Yl, P2 = Xr*Yr, P3 = (Xl+Xr)*(Yl+Yr),
//   X*Y = P1 + 2^h (P3 - P1 - P2) + 2^2h P2
// as: one vector addition (the two sums), ONE vector multiplication of the three half-size
// products, and three additions.  Full 2*nbits-bit product.  (The reference truncates the
// sums to h bits and concatenates the pieces without carries, so it is only right for small
// operands; here the sums keep their carry bit and the recombination is a real addition.)
tfhe_b200_circuit *tfhe_b200_circuit_mul_karatsuba(tfhe_b200_ctx *ctx, int nbits, int count, int adder) {
    if (nbits < 2 || (nbits & 1) || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    const int W = 2 * nbits, h = nbits / 2, e = h + 1;
    c->out_row0 = B.alloc(count * W);
    c->out_rows = count * W;
    c->row_zero = B.alloc(1);
    const int Z = c->row_zero;
    std::vector<Bits> lo, hi, sums;  // [Xl.., Yl..], [Xr.., Yr..], [Sx.., Sy..], each e bits (zero extended)
    for (int side = 0; side < 2; side++)
        for (int i = 0; i < count; i++) {
            const int base = (side == 0 ? a : b) + i * nbits;
            Bits l = Builder::bits_at(base, h), r = Builder::bits_at(base + h, h);
            l.push_back(Z);
            r.push_back(Z);
            lo.push_back(l);
            hi.push_back(r);
            sums.push_back(B.alloc_bits(e));
        }
    B.add(adder, lo, hi, sums, e);
    // the three products of e-bit operands, 2e bits each: [P1.., P2.., P3..]
    std::vector<Bits> ma, mb, mo;
    for (int which = 0; which < 3; which++)
        for (int i = 0; i < count; i++) {
            const std::vector<Bits> &src = which == 0 ? lo : (which == 1 ? hi : sums);
            ma.push_back(src[i]);
            mb.push_back(src[count + i]);
            mo.push_back(B.alloc_bits(2 * e));
        }
    build_mul_full(B, ma, mb, mo, e, adder, Z);
    // E = P3 - (P1 + P2)
    std::vector<Bits> p1(mo.begin(), mo.begin() + count), p2(mo.begin() + count, mo.begin() + 2 * count),
        p3(mo.begin() + 2 * count, mo.end()), t(count), E(count);
    for (int i = 0; i < count; i++) {
        t[i] = B.alloc_bits(2 * e);
        E[i] = B.alloc_bits(2 * e);
    }
    B.add(adder, p1, p2, t, 2 * e);
    B.sub(adder, p3, t, E, 2 * e);
    // result = (P1 | P2 << 2h) + (E << h): bits below h are P1's, the rest is one addition
    std::vector<Bits> base_hi(count), e_hi(count), out_hi(count);
    for (int i = 0; i < count; i++) {
        for (int k = 0; k < h; k++) B.copy(p1[i][k], c->out_row0 + i * W + k);
        for (int k = h; k < W; k++) {
            base_hi[i].push_back(k < 2 * h ? p1[i][k] : p2[i][k - 2 * h]);
            e_hi[i].push_back(k - h < 2 * e ? E[i][k - h] : Z);
            out_hi[i].push_back(c->out_row0 + i * W + k);
        }
    }
    B.add(adder, base_hi, e_hi, out_hi, W - h);
    B.end_level();
    return finish(c);
}

// Cannon's algorithm (BOOTS_CannonsAlgo, main.cu:2590-2644) for square n x n matrices: n steps of
// one vector multiplication of the n*n aligned pairs (A[i][(i+j+s) % n], B[(i+j+s) % n][j]) and one
// vector addition into C.  The reference rotates rows / columns between the steps; here a step is
// just another index table.  Same result as tfhe_b200_circuit_matmul with a working set of n*n
// instead of n*n*n products.
tfhe_b200_circuit *tfhe_b200_circuit_matmul_cannon(tfhe_b200_ctx *ctx, int n, int nbits, int adder) {
    if (n < 1 || nbits < 2 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int A = B.operand(n * n * nbits), Bm = B.operand(n * n * nbits);
    c->out_row0 = B.alloc(n * n * nbits);
    c->out_rows = n * n * nbits;
    c->row_zero = -2;
    std::vector<Bits> acc;
    for (int s = 0; s < n; s++) {
        std::vector<Bits> va, vb, prod, sum;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                const int k = (i + j + s) % n;
                va.push_back(Builder::bits_at(A + (i * n + k) * nbits, nbits));
                vb.push_back(Builder::bits_at(Bm + (k * n + j) * nbits, nbits));
                const Bits o = Builder::bits_at(c->out_row0 + (i * n + j) * nbits, nbits);
                prod.push_back((n == 1) ? o : B.alloc_bits(nbits));
                sum.push_back(s == n - 1 ? o : B.alloc_bits(nbits));
            }
        build_mul(B, va, vb, prod, nbits, adder);
        if (s == 0) {
            acc = prod;
        } else {
            B.add(adder, acc, prod, sum, nbits);
            acc = sum;
        }
    }
    return finish(c);
}

// ---- the rest of the Cipher arithmetic (Cipher.cu:237-630) as plans ------------------------

// a - b mod 2^nbits (operator-, Cipher.cu:329-332)
tfhe_b200_circuit *tfhe_b200_circuit_sub(tfhe_b200_ctx *ctx, int nbits, int count, int adder) {
    if (nbits < 1 || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    B.sub(adder, numbers_at(a, count, nbits), numbers_at(b, count, nbits), numbers_at(c->out_row0, count, nbits), nbits);
    return finish(c);
}

// -a mod 2^nbits (twosComplement, Cipher.cu:286-298)
tfhe_b200_circuit *tfhe_b200_circuit_neg(tfhe_b200_ctx *ctx, int nbits, int count) {
    if (nbits < 1 || count < 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    B.negate(numbers_at(a, count, nbits), numbers_at(c->out_row0, count, nbits), nbits);
    return finish(c);
}

// One result bit per pair: op = TFHE_B200_CMP_* (operator>, operator<=, operator== of Cipher.cu:561-616
// and their complements).
tfhe_b200_circuit *tfhe_b200_circuit_compare(tfhe_b200_ctx *ctx, int nbits, int count, int op, int is_signed) {
    if (nbits < 1 || count < 1 || op < 0 || op > TFHE_B200_CMP_NE) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count);
    c->out_rows = count;
    const std::vector<Bits> va = numbers_at(a, count, nbits), vb = numbers_at(b, count, nbits);
    std::vector<int> r;
    bool invert = false;
    switch (op) {
        case TFHE_B200_CMP_GT: r = B.greater(va, vb, nbits, is_signed != 0); break;
        case TFHE_B200_CMP_LE: r = B.greater(va, vb, nbits, is_signed != 0); invert = true; break;
        case TFHE_B200_CMP_LT: r = B.greater(vb, va, nbits, is_signed != 0); break;
        case TFHE_B200_CMP_GE: r = B.greater(vb, va, nbits, is_signed != 0); invert = true; break;
        case TFHE_B200_CMP_EQ: r = B.equal(va, vb, nbits); break;
        default: r = B.equal(va, vb, nbits); invert = true; break;
    }
    for (int i = 0; i < count; i++) {
        if (invert) B.not_(r[i], c->out_row0 + i);
        else B.copy(r[i], c->out_row0 + i);
    }
    B.end_level();
    return finish(c);
}

// min(a, b) / max(a, b) (minimum, Cipher.cu:301-320: comparison, then one bootsMUX per bit)
tfhe_b200_circuit *tfhe_b200_circuit_minmax(tfhe_b200_ctx *ctx, int nbits, int count, int want_max, int is_signed) {
    if (nbits < 1 || count < 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    const std::vector<Bits> va = numbers_at(a, count, nbits), vb = numbers_at(b, count, nbits);
    const std::vector<int> gt = B.greater(va, vb, nbits, is_signed != 0);
    for (int i = 0; i < count; i++)
        for (int bit = 0; bit < nbits; bit++) {
            const int out = c->out_row0 + i * nbits + bit;
            if (want_max) B.mux(gt[i], va[i][bit], vb[i][bit], out);
            else B.mux(gt[i], vb[i][bit], va[i][bit], out);
        }
    B.end_level();
    return finish(c);
}

// sel ? a : b for count numbers (one selector bit per number; bootsMUX per bit, Cipher::mux)
tfhe_b200_circuit *tfhe_b200_circuit_select(tfhe_b200_ctx *ctx, int nbits, int count) {
    if (nbits < 1 || count < 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int sel = B.operand(count), a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    for (int i = 0; i < count; i++)
        for (int bit = 0; bit < nbits; bit++)
            B.mux(sel + i, a + i * nbits + bit, b + i * nbits + bit, c->out_row0 + i * nbits + bit);
    B.end_level();
    return finish(c);
}

// |a| for two's complement a (absolute, Cipher.cu:469-492: mask = sign bit replicated,
// (a + mask) ^ mask)
static void build_abs(Builder &B, const std::vector<Bits> &a, const std::vector<Bits> &out, int nbits, int adder) {
    const int m = (int) a.size();
    std::vector<Bits> mask(m), sum(m);
    for (int i = 0; i < m; i++) {
        mask[i].assign(nbits, a[i][nbits - 1]);  // every bit of the mask IS the sign row
        sum[i] = B.alloc_bits(nbits);
    }
    B.add(adder, a, mask, sum, nbits);
    for (int i = 0; i < m; i++)
        for (int bit = 0; bit < nbits; bit++) B.gate(TFHE_B200_XOR, sum[i][bit], a[i][nbits - 1], out[i][bit]);
    B.end_level();
}

tfhe_b200_circuit *tfhe_b200_circuit_abs(tfhe_b200_ctx *ctx, int nbits, int count, int adder) {
    if (nbits < 2 || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    build_abs(B, numbers_at(a, count, nbits), numbers_at(c->out_row0, count, nbits), nbits, adder);
    return finish(c);
}

// Shifts by a public amount: kind = TFHE_B200_SHIFT_LEFT (innerLeftShift, Cipher.cu:215-228: zeros
// come in), _RIGHT_LOGICAL, _RIGHT_ARITH (sign fill: the copy part of rightShift, Cipher.cu:237-246).
// Bootstrap free.
tfhe_b200_circuit *tfhe_b200_circuit_shift(tfhe_b200_ctx *ctx, int nbits, int count, int amount, int kind) {
    if (nbits < 1 || count < 1 || amount < 0 || kind < 0 || kind > TFHE_B200_SHIFT_RIGHT_ARITH) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * nbits);
    c->out_rows = count * nbits;
    for (int i = 0; i < count; i++)
        for (int bit = 0; bit < nbits; bit++) {
            const int out = c->out_row0 + i * nbits + bit;
            const int src = (kind == TFHE_B200_SHIFT_LEFT) ? bit - amount : bit + amount;
            if (src >= 0 && src < nbits) B.copy(a + i * nbits + src, out);
            else if (kind == TFHE_B200_SHIFT_RIGHT_ARITH) B.copy(a + i * nbits + nbits - 1, out);
            else B.constant(0, out);
        }
    B.end_level();
    return finish(c);
}

// Division (operator/, divInternal, addSign; Cipher.cu:494-559): restoring division of the
// absolute values, nbits rounds of shift / subtract / select, then the sign of the quotient.
// Output per pair: quotient (nbits rows) followed by remainder (nbits rows; the remainder of
// |a| / |b|, as divInternal leaves it).  is_signed = 0: operands are unsigned, no sign handling.
// Division by zero gives the all-ones quotient pattern of the restoring algorithm (reference too).
tfhe_b200_circuit *tfhe_b200_circuit_div(tfhe_b200_ctx *ctx, int nbits, int count, int is_signed, int adder) {
    if (nbits < 2 || count < 1 || adder < 0 || adder > 1) return nullptr;
    tfhe_b200_circuit *c = new_plan(ctx);
    Builder B(c);
    const int a = B.operand(count * nbits), b = B.operand(count * nbits);
    c->out_row0 = B.alloc(count * 2 * nbits);
    c->out_rows = count * 2 * nbits;
    c->row_zero = B.alloc(1);
    std::vector<Bits> ua = numbers_at(a, count, nbits), ub = numbers_at(b, count, nbits);
    if (is_signed) {
        std::vector<Bits> both, both_out;
        for (int i = 0; i < count; i++) {
            both.push_back(ua[i]);
            both_out.push_back(B.alloc_bits(nbits));
        }
        for (int i = 0; i < count; i++) {
            both.push_back(ub[i]);
            both_out.push_back(B.alloc_bits(nbits));
        }
        build_abs(B, both, both_out, nbits, adder);
        for (int i = 0; i < count; i++) {
            ua[i] = both_out[i];
            ub[i] = both_out[count + i];
        }
    }
    // the subtraction runs on nbits + 1 bits (P can exceed nbits bits by one before the restore)
    const int w = nbits + 1;
    std::vector<Bits> P(count), A = ua, Bx(count);  // P: partial remainder (w bits), A: dividend bits still to shift in
    for (int i = 0; i < count; i++) {
        P[i].assign(w, c->row_zero);
        Bx[i] = ub[i];
        Bx[i].push_back(c->row_zero);  // zero-extended divisor
    }
    std::vector<Bits> quot(count, Bits(nbits, -1));
    for (int step = 0; step < nbits; step++) {
        // PA <<= 1 : P takes the top dividend bit (index renaming only)
        std::vector<Bits> Ps(count), diff(count);
        for (int i = 0; i < count; i++) {
            Ps[i].resize(w);
            Ps[i][0] = A[i][nbits - 1 - step];
            for (int k = 1; k < w; k++) Ps[i][k] = P[i][k - 1];
            diff[i] = B.alloc_bits(w);
        }
        B.sub(adder, Ps, Bx, diff, w);
        // quotient bit = NOT sign(diff); P = sign ? Ps : diff   (divInternal, Cipher.cu:524-541)
        for (int i = 0; i < count; i++) {
            quot[i][nbits - 1 - step] = B.alloc(1);
            B.not_(diff[i][w - 1], quot[i][nbits - 1 - step]);
            Bits np = B.alloc_bits(w);
            for (int k = 0; k < w; k++) B.mux(diff[i][w - 1], Ps[i][k], diff[i][k], np[k]);
            P[i] = np;
        }
        B.end_level();
    }
    if (is_signed) {
        // addSign (Cipher.cu:543-559): negate the quotient when the operand signs differ
        std::vector<Bits> nq(count);
        std::vector<int> sgn(count);
        for (int i = 0; i < count; i++) {
            nq[i] = B.alloc_bits(nbits);
            sgn[i] = B.alloc(1);
            B.gate(TFHE_B200_XOR, a + i * nbits + nbits - 1, b + i * nbits + nbits - 1, sgn[i]);
        }
        B.end_level();
        B.negate(quot, nq, nbits);
        for (int i = 0; i < count; i++)
            for (int k = 0; k < nbits; k++) B.mux(sgn[i], nq[i][k], quot[i][k], c->out_row0 + i * 2 * nbits + k);
    } else {
        for (int i = 0; i < count; i++)
            for (int k = 0; k < nbits; k++) B.copy(quot[i][k], c->out_row0 + i * 2 * nbits + k);
    }
    for (int i = 0; i < count; i++)
        for (int k = 0; k < nbits; k++) B.copy(P[i][k], c->out_row0 + i * 2 * nbits + nbits + k);
    B.end_level();
    return finish(c);
}

void tfhe_b200_circuit_destroy(tfhe_b200_circuit *c) {
    if (!c) return;
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->d_idx) cudaFree(c->d_idx);
    if (c->d_ws) cudaFree(c->d_ws);
    if (c->d_u) cudaFree(c->d_u);
    delete c;
}

int tfhe_b200_circuit_levels(const tfhe_b200_circuit *c) { return c ? c->depth : 0; }
// bootstraps of the level-th sequential bootstrap batch (0 .. levels - 1); a MUX group counts two per row
long long tfhe_b200_circuit_level_gates(const tfhe_b200_circuit *c, int level) {
    if (!c || level < 0) return 0;
    for (const Level &lv : c->levels) {
        if (lv.type == LV_LINEAR) continue;
        if (level-- > 0) continue;
        long long n = 0;
        for (int i = 0; i < lv.nops; i++) n += lv.ops[i].count;
        return lv.type == LV_MUX ? 2 * n : n;
    }
    return 0;
}
long long tfhe_b200_circuit_gates(const tfhe_b200_circuit *c) { return c ? c->n_gates : 0; }
int tfhe_b200_circuit_operands(const tfhe_b200_circuit *c) { return c ? (int) c->in_row0.size() : 0; }
int tfhe_b200_circuit_operand_rows(const tfhe_b200_circuit *c, int o) {
    return (c && o >= 0 && o < (int) c->in_rows.size()) ? c->in_rows[o] : 0;
}
int tfhe_b200_circuit_output_rows(const tfhe_b200_circuit *c) { return c ? c->out_rows : 0; }

}  // extern "C"

namespace {

// one run of a GATES group as the engine's gate_op (rows of the plan's workspace, index tables)
void fill_gate_op(const tfhe_b200_circuit *c, const Op &op, tfhe_b200_gate_op &o) {
    o.gate = op.gate;
    o.count = op.count;
    o.a = o.b = c->d_ws;
    o.out = c->d_ws;
    o.stride_a = o.stride_b = o.stride_out = c->words;
    o.idx_a = c->d_idx + op.off_a;
    o.idx_b = c->d_idx + op.off_b;
    o.idx_out = c->d_idx + op.off_out;
    o.c = op.off_c >= 0 ? c->d_ws : nullptr;
    o.stride_c = c->words;
    o.idx_c = op.off_c >= 0 ? c->d_idx + op.off_c : nullptr;
    o.d = op.off_d >= 0 ? c->d_ws : nullptr;
    o.stride_d = c->words;
    o.idx_d = op.off_d >= 0 ? c->d_idx + op.off_d : nullptr;
}

int issue_mux_or_linear(tfhe_b200_circuit *c, const Level &lv, void *stream) {
    const Op &op = lv.ops[0];
    if (lv.type == LV_MUX)
        return tfhe_b200_mux_gather(c->ctx, c->d_ws, c->words, c->d_idx + op.off_a, c->d_idx + op.off_b,
                                    c->d_idx + op.off_c, c->d_idx + op.off_out, op.count, stream);
    const int coef = op.gate == LIN_COPY ? 1 : (op.gate == LIN_NOT ? -1 : 0);
    const int32_t cst = op.gate == LIN_ONE ? 0x20000000 : (op.gate == LIN_ZERO ? (int32_t) 0xE0000000 : 0);
    return tfhe_b200_linear_gather(c->ctx, c->d_ws, c->words, c->d_idx + op.off_a, c->d_idx + op.off_out, coef, cst,
                                   op.count, stream);
}

// every launch group of the plan, in order, on `stream` (scratch: the plan's own)
int issue_levels(tfhe_b200_circuit *c, void *stream) {
    tfhe_b200::engine_set_thread_scratch(c->d_u, c->u_bytes);
    int rc = 0;
    for (const Level &lv : c->levels) {
        if (lv.type == LV_GATES) {
            tfhe_b200_gate_op ops[4];
            for (int i = 0; i < lv.nops; i++) fill_gate_op(c, lv.ops[i], ops[i]);
            rc = tfhe_b200_gate_multi(c->ctx, ops, lv.nops, stream);
        } else {
            rc = issue_mux_or_linear(c, lv, stream);
        }
        if (rc) break;
    }
    tfhe_b200::engine_set_thread_scratch(nullptr, 0);
    return rc;
}

int copy_operands_in(tfhe_b200_circuit *c, const int32_t *const *operands, void *stream) {
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
    return 0;
}

int copy_result_out(tfhe_b200_circuit *c, int32_t *d_out, void *stream) {
    const size_t rb = (size_t) c->words * sizeof(int32_t);
    if (cudaMemcpyAsync(d_out, c->d_ws + (size_t) c->out_row0 * c->words, rb * c->out_rows, cudaMemcpyDeviceToDevice,
                        (cudaStream_t) stream) != cudaSuccess)
        return fail_msg("result copy failed");
    return 0;
}

// Capture the launch sequence once (thread-local capture mode: other threads keep using the device).
// A failed capture is not an error: the plan then launches its kernels directly, as before.
void try_capture(tfhe_b200_circuit *c, cudaStream_t st) {
    const unsigned long long l0 = tfhe_b200_launch_count(c->ctx);
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->graph_failed = true;
        return;
    }
    const int rc = issue_levels(c, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const unsigned long long l1 = tfhe_b200_launch_count(c->ctx);
    if (rc == 0 && e == cudaSuccess && graph != nullptr &&
        cudaGraphInstantiate(&c->graph_exec, graph, nullptr, nullptr, 0) == cudaSuccess) {
        c->graph_launches = l1 - l0;
    } else {
        cudaGetLastError();
        c->graph_exec = nullptr;
        c->graph_failed = true;
    }
    if (graph) cudaGraphDestroy(graph);
    // the capture itself launched nothing: take its count back
    tfhe_b200_count_launches(c->ctx, (unsigned long long) 0 - (l1 - l0));
}

}  // namespace

extern "C" {

int tfhe_b200_circuit_set_graph(tfhe_b200_circuit *c, int enable) {
    if (!c) return 0;
    const int prev = c->graph_enabled ? 1 : 0;
    c->graph_enabled = enable != 0;
    return prev;
}

int tfhe_b200_circuit_used_graph(const tfhe_b200_circuit *c) { return (c && c->used_graph) ? 1 : 0; }

// Runs the plan: operands[o] and d_out are DEVICE arrays of samples (rows of n+1 words).
int tfhe_b200_circuit_run(tfhe_b200_circuit *c, int32_t *d_out, const int32_t *const *operands, void *stream) {
    if (!c || !d_out || !operands) return fail_msg("null argument");
    if (!c->ctx) return fail_msg("plan was built without an engine context (simulation only)");
    if (ensure_device(c)) return 1;
    cudaStream_t st = (cudaStream_t) stream;
    if (cudaSetDevice(tfhe_b200_ctx_device(c->ctx)) != cudaSuccess) return fail_msg("cudaSetDevice failed");
    if (copy_operands_in(c, operands, stream)) return 1;
    c->used_graph = false;
    // the legacy default stream cannot be captured; plans of fewer than three groups gain nothing
    if (c->graph_enabled && !c->graph_failed && st != nullptr && c->graph_exec == nullptr && c->levels.size() >= 3)
        try_capture(c, st);
    if (c->graph_enabled && c->graph_exec != nullptr && st != nullptr) {
        if (cudaGraphLaunch(c->graph_exec, st) != cudaSuccess) return fail_msg("graph launch failed");
        tfhe_b200_count_launches(c->ctx, c->graph_launches);
        c->used_graph = true;
    } else if (issue_levels(c, stream)) {
        return 1;
    }
    return copy_result_out(c, d_out, stream);
}

// K independent plans, zipped group by group: the GATES groups at the same position of all plans share
// launches (up to TFHE_B200_MAX_RUNS runs each); MUX and LINEAR groups are issued per plan.
int tfhe_b200_circuit_run_many(tfhe_b200_circuit *const *plans, int nplans, int32_t *const *d_outs,
                               const int32_t *const *const *operands, void *stream) {
    if (!plans || !d_outs || !operands || nplans < 1) return fail_msg("null argument");
    tfhe_b200_ctx *ctx = plans[0] ? plans[0]->ctx : nullptr;
    if (!ctx) return fail_msg("plan was built without an engine context (simulation only)");
    size_t longest = 0, u_need = 0;
    for (int p = 0; p < nplans; p++) {
        tfhe_b200_circuit *c = plans[p];
        if (!c || c->ctx != ctx) return fail_msg("all plans of a merged run must belong to one context");
        for (int q = 0; q < p; q++)
            if (plans[q] == c) return fail_msg("a plan may appear only once in a merged run (it has one workspace)");
        if (ensure_device(c)) return 1;
        if (c->levels.size() > longest) longest = c->levels.size();
    }
    if (cudaSetDevice(tfhe_b200_ctx_device(ctx)) != cudaSuccess) return fail_msg("cudaSetDevice failed");
    // scratch of a merged launch: the sum of the plans' widest levels bounds every merged level
    for (int p = 0; p < nplans; p++) u_need += plans[p]->u_bytes;
    int32_t *d_u = nullptr;
    if (cudaMallocAsync(&d_u, u_need, (cudaStream_t) stream) != cudaSuccess) return fail_msg("scratch allocation failed");
    int rc = 0;
    for (int p = 0; p < nplans && !rc; p++) rc = copy_operands_in(plans[p], operands[p], stream);
    tfhe_b200::engine_set_thread_scratch(d_u, u_need);
    for (size_t pos = 0; pos < longest && !rc; pos++) {
        tfhe_b200_gate_op ops[TFHE_B200_MAX_RUNS];
        int nops = 0;
        for (int p = 0; p < nplans && !rc; p++) {
            tfhe_b200_circuit *c = plans[p];
            if (pos >= c->levels.size()) continue;
            const Level &lv = c->levels[pos];
            if (lv.type != LV_GATES) continue;
            if (nops + lv.nops > TFHE_B200_MAX_RUNS) {
                rc = tfhe_b200_gate_multi(ctx, ops, nops, stream);
                nops = 0;
            }
            for (int i = 0; i < lv.nops; i++) fill_gate_op(c, lv.ops[i], ops[nops++]);
        }
        if (!rc && nops) rc = tfhe_b200_gate_multi(ctx, ops, nops, stream);
        for (int p = 0; p < nplans && !rc; p++) {
            tfhe_b200_circuit *c = plans[p];
            if (pos < c->levels.size() && c->levels[pos].type != LV_GATES) rc = issue_mux_or_linear(c, c->levels[pos], stream);
        }
    }
    tfhe_b200::engine_set_thread_scratch(nullptr, 0);
    for (int p = 0; p < nplans && !rc; p++) rc = copy_result_out(plans[p], d_outs[p], stream);
    cudaFreeAsync(d_u, (cudaStream_t) stream);
    return rc;
}

// Evaluates the plan on PLAINTEXT bits on the host (one int per sample row), group by group
// in the order the device runs them: the schedule's logic can be checked without keys or a GPU.
// operand_bits[o]: operand_rows(o) ints in {0,1}.
int tfhe_b200_circuit_simulate(const tfhe_b200_circuit *c, int32_t *out_bits, const int32_t *const *operand_bits) {
    if (!c || !out_bits || !operand_bits) return fail_msg("null argument");
    std::vector<int32_t> ws((size_t) c->nrows, 0);
    for (size_t o = 0; o < c->in_row0.size(); o++)
        for (int r = 0; r < c->in_rows[o]; r++) ws[c->in_row0[o] + r] = operand_bits[o][r] & 1;
    for (const Level &lv : c->levels) {
        // all gates of a group read the state before the group (they run as one batch)
        std::vector<std::pair<int, int32_t>> writes;
        for (int i = 0; i < lv.nops; i++) {
            const Op &op = lv.ops[i];
            for (int g = 0; g < op.count; g++) {
                int v;
                if (lv.type == LV_GATES) {
                    const int a = ws[c->h_idx[op.off_a + g]], b = ws[c->h_idx[op.off_b + g]];
                    const int c3 = op.off_c >= 0 ? ws[c->h_idx[op.off_c + g]] : 0;
                    const int c4 = op.off_d >= 0 ? ws[c->h_idx[op.off_d + g]] : 0;
                    v = gate_truth(op.gate, a, b, c3, c4);
                } else if (lv.type == LV_MUX) {
                    v = ws[c->h_idx[op.off_a + g]] ? ws[c->h_idx[op.off_b + g]] : ws[c->h_idx[op.off_c + g]];
                } else {
                    const int in = ws[c->h_idx[op.off_a + g]];
                    v = op.gate == LIN_COPY ? in : (op.gate == LIN_NOT ? !in : (op.gate == LIN_ONE ? 1 : 0));
                }
                writes.emplace_back(c->h_idx[op.off_out + g], v);
            }
        }
        for (const auto &w : writes) ws[w.first] = w.second;
    }
    for (int r = 0; r < c->out_rows; r++) out_bits[r] = ws[c->out_row0 + r];
    return 0;
}

}  // extern "C"
