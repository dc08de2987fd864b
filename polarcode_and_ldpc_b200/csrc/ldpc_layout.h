// ldpc_layout.h -- host-side construction of the conflict-free shared-memory layout used by
// ldpc_banked_kernel (see ldpc_banked.cuh for the layout itself).  Plain C++, no CUDA.
//
// Input: a regular code, every variable in exactly 3 checks, every check of degree dc.
// Output: a seat (round R, lane) for every check, a position (round r, lane) for every
// variable, and for every variable the order in which its three messages are fetched, such that
// the 32 words one warp instruction of the variable pass touches lie in distinct banks
// (bank = lane of the check).  A round's 96 edges have to hit every bank exactly three times,
// which a short simulated-annealing run over (check -> lane, variable -> round) gets to within
// a handful of edges; the per-round assignment of edges to the three instructions is then a
// bipartite edge colouring, found by local search over the 6 orders of each variable.
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

struct BankedLayout {
    int nR = 0, NP = 0, NS = 0;
    int residual = 0;                       // edges that still share a bank with another one of their instruction
    std::vector<unsigned long long> bpack;  // [NP] three slots, 16 bits each; bits 48.. set for an empty position
    std::vector<uint16_t> varof;            // [NP] variable at a position (0xffff: empty)
    std::vector<uint16_t> posof;            // [n]
    std::vector<uint16_t> cpos;             // [NS] variable position behind a slot (0xffff: empty)
};

struct LayoutRng {
    unsigned long long s;
    explicit LayoutRng(unsigned long long seed) : s(seed * 0x9E3779B97F4A7C15ull + 88172645463325252ull) {}
    unsigned next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (unsigned)(s >> 11); }
    double u01() { return (next() & 0xffffff) / 16777216.0; }
};

// var_checks[v][0..2] ascending, check_vars[c][0..dc-1] ascending
// paired = true: the k-th message of the check at (R, lane) is word 64 (dc/2 R + k/2) + 2 lane + (k & 1),
// so a check moves its messages with 8-byte accesses and an edge's bank is 2 (lane % 16) + (k & 1):
// the search also chooses, per check, which half of its edges sit at odd k.
static bool build_banked_layout(int m, int n, int dc, const std::vector<int>& var_checks,
                                const std::vector<int>& check_vars, long budget, BankedLayout* out, bool paired = false)
{
    const int nr = (n + 31) / 32, nR = (m + 31) / 32;
    if (paired && (dc & 1)) return false;
    // par[c * dc + k]: parity of the seat of the k-th (ascending) variable of check c
    std::vector<unsigned char> par((size_t)m * dc, 0);
    if (paired)
        for (int c = 0; c < m; c++)
            for (int k = 0; k < dc; k++) par[(size_t)c * dc + k] = (unsigned char)(k & 1);
    auto edge_k = [&](int c, int v) {
        int k = 0;
        while (check_vars[(size_t)c * dc + k] != v) k++;
        return k;
    };
    auto bank_of = [&](int c, int k, int lane) { return paired ? 2 * (lane & 15) + par[(size_t)c * dc + k] : lane; };
    if ((long)32 * dc * nR > 65535 || 32 * nr > 65535) return false;
    std::vector<int> lam(m), rd(n), lamcnt(32, 0), rcnt(nr, 0), C((size_t)nr * 32, 0);
    for (int c = 0; c < m; c++) { lam[c] = c % 32; lamcnt[lam[c]]++; }
    for (int v = 0; v < n; v++) { rd[v] = v / 32; rcnt[rd[v]]++; }
    // vk[3 v + j]: index of variable v among the variables of its j-th check
    std::vector<int> vk(3 * (size_t)n);
    for (int v = 0; v < n; v++)
        for (int j = 0; j < 3; j++) vk[3 * v + j] = edge_k(var_checks[3 * v + j], v);
    for (int v = 0; v < n; v++)
        for (int j = 0; j < 3; j++) {
            const int c = var_checks[3 * v + j];
            C[(size_t)rd[v] * 32 + bank_of(c, vk[3 * v + j], lam[c])]++;
        }
    auto ex = [](int x) { return x > 3 ? x - 3 : 0; };
    long cur = 0;
    for (int x : C) cur += ex(x);
    LayoutRng rng(0x5eed0000ull + (unsigned long long)n * 131 + m);
    double T = 0.2;
    const long cool = budget / 130 + 1;
    for (long it = 0; it < budget && cur > 0; it++) {
        const unsigned mv = rng.next() % 10;
        if (paired && mv == 9) {
            // swap the seat parity of an even-seated and an odd-seated edge of one check
            const int c = rng.next() % m;
            int ke = rng.next() % dc, ko = rng.next() % dc;
            if (par[(size_t)c * dc + ke] != 0 || par[(size_t)c * dc + ko] != 1) continue;
            const int re = rd[check_vars[(size_t)c * dc + ke]], ro = rd[check_vars[(size_t)c * dc + ko]];
            const int b0 = 2 * (lam[c] & 15), b1 = b0 + 1;
            long d = 0;
            int* cell[4] = {&C[(size_t)re * 32 + b0], &C[(size_t)re * 32 + b1], &C[(size_t)ro * 32 + b1], &C[(size_t)ro * 32 + b0]};
            const int dv[4] = {-1, 1, -1, 1};
            for (int t = 0; t < 4; t++) { d += ex(*cell[t] + dv[t]) - ex(*cell[t]); *cell[t] += dv[t]; }
            if (d <= 0 || rng.u01() < exp(-(double)d / T)) {
                par[(size_t)c * dc + ke] = 1; par[(size_t)c * dc + ko] = 0;
                cur += d;
            } else {
                for (int t = 0; t < 4; t++) *cell[t] -= dv[t];
            }
        } else if (mv < 5) {
            // move a check to another lane (swap with a check of that lane when the lane is full)
            const int a = rng.next() % m, lb = rng.next() % 32, la = lam[a];
            if (la == lb) continue;
            int b = -1;
            if (lamcnt[lb] >= nR) {
                for (int t = 0; t < 8 && b < 0; t++) { const int cnd = rng.next() % m; if (lam[cnd] == lb) b = cnd; }
                if (b < 0) continue;
            }
            long d = 0;
            auto shift = [&](int c, int from, int to, int sgn) {
                for (int k = 0; k < dc; k++) {
                    const int r = rd[check_vars[(size_t)c * dc + k]];
                    int& x = C[(size_t)r * 32 + bank_of(c, k, from)];
                    int& y = C[(size_t)r * 32 + bank_of(c, k, to)];
                    if (&x == &y) continue;
                    if (sgn > 0) { d += ex(x - 1) - ex(x); x--; d += ex(y + 1) - ex(y); y++; }
                    else { x++; y--; }
                }
            };
            shift(a, la, lb, 1);
            if (b >= 0) shift(b, lb, la, 1);
            if (d <= 0 || rng.u01() < exp(-(double)d / T)) {
                lam[a] = lb; lamcnt[la]--; lamcnt[lb]++;
                if (b >= 0) { lam[b] = la; lamcnt[lb]--; lamcnt[la]++; }
                cur += d;
            } else {
                shift(a, la, lb, -1);
                if (b >= 0) shift(b, lb, la, -1);
            }
        } else {
            // move a variable to another round (swap when the round is full)
            const int a = rng.next() % n, rb = rng.next() % nr, ra = rd[a];
            if (ra == rb) continue;
            int b = -1;
            if (rcnt[rb] >= 32) {
                for (int t = 0; t < 8 && b < 0; t++) { const int cnd = rng.next() % n; if (rd[cnd] == rb) b = cnd; }
                if (b < 0) continue;
            }
            long d = 0;
            auto shift = [&](int v, int from, int to, int sgn) {
                for (int j = 0; j < 3; j++) {
                    const int l = bank_of(var_checks[3 * v + j], vk[3 * v + j], lam[var_checks[3 * v + j]]);
                    int& x = C[(size_t)from * 32 + l];
                    int& y = C[(size_t)to * 32 + l];
                    if (sgn > 0) { d += ex(x - 1) - ex(x); x--; d += ex(y + 1) - ex(y); y++; }
                    else { x++; y--; }
                }
            };
            shift(a, ra, rb, 1);
            if (b >= 0) shift(b, rb, ra, 1);
            if (d <= 0 || rng.u01() < exp(-(double)d / T)) {
                rd[a] = rb; rcnt[ra]--; rcnt[rb]++;
                if (b >= 0) { rd[b] = ra; rcnt[rb]--; rcnt[ra]++; }
                cur += d;
            } else {
                shift(a, ra, rb, -1);
                if (b >= 0) shift(b, rb, ra, -1);
            }
        }
        if (it % cool == 0) { T *= 0.97; if (T < 0.03) T = 0.03; }
    }

    // seats: the checks of a lane take rounds 0, 1, ..; the variables of a round take lanes 0, 1, ..
    std::vector<int> seatR(m), fill(32, 0), vlane(n), vfill(nr, 0);
    for (int c = 0; c < m; c++) seatR[c] = fill[lam[c]]++;
    for (int v = 0; v < n; v++) vlane[v] = vfill[rd[v]]++;
    out->nR = nR; out->NP = 32 * nr; out->NS = 32 * dc * nR;
    out->varof.assign(out->NP, 0xffff);
    out->posof.assign(n, 0);
    out->cpos.assign(out->NS, 0xffff);
    out->bpack.assign(out->NP, 0xffffull << 48);
    for (int v = 0; v < n; v++) {
        const int pi = 32 * rd[v] + vlane[v];
        out->varof[pi] = (uint16_t)v;
        out->posof[v] = (uint16_t)pi;
    }
    // seat index kk of every edge inside its check: ascending variables (paired: the even-parity
    // edges take kk = 0, 2, .., the odd ones 1, 3, ..); slot = 32 (dc R + kk) + lane, or
    // 64 (dc/2 R + kk/2) + 2 lane + (kk & 1) in the paired layout
    std::vector<int> seatk((size_t)m * dc, 0);
    for (int c = 0; c < m; c++) {
        int ne = 0, no = 0;
        for (int k = 0; k < dc; k++) {
            if (!paired) seatk[(size_t)c * dc + k] = k;
            else if (par[(size_t)c * dc + k]) seatk[(size_t)c * dc + k] = 2 * (no++) + 1;
            else seatk[(size_t)c * dc + k] = 2 * (ne++);
        }
        if (paired && (ne != dc / 2 || no != dc / 2)) return false;
    }
    auto slot_ck = [&](int c, int k) {
        const int kk = seatk[(size_t)c * dc + k];
        return paired ? 64 * ((dc / 2) * seatR[c] + (kk >> 1)) + 2 * lam[c] + (kk & 1) : 32 * (dc * seatR[c] + kk) + lam[c];
    };
    auto slot_of = [&](int c, int v) { return slot_ck(c, edge_k(c, v)); };
    for (int c = 0; c < m; c++)
        for (int k = 0; k < dc; k++) out->cpos[slot_ck(c, k)] = out->posof[check_vars[(size_t)c * dc + k]];

    // per round: which of the three instructions fetches which edge (local search over the 6 orders)
    static const int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    std::vector<int> order(n, 0);
    int residual = 0;
    for (int r = 0; r < nr; r++) {
        std::vector<int> vs;
        for (int v = 0; v < n; v++)
            if (rd[v] == r) vs.push_back(v);
        int mult[3][32] = {};
        auto add = [&](int v, int pidx, int sgn) {
            for (int j = 0; j < 3; j++) {
                const int jj = PERM[pidx][j], c = var_checks[3 * v + jj];
                mult[j][bank_of(c, vk[3 * v + jj], lam[c])] += sgn;
            }
        };
        auto cost = [&]() {
            int cst = 0;
            for (int j = 0; j < 3; j++)
                for (int l = 0; l < 32; l++) cst += mult[j][l] > 1 ? mult[j][l] - 1 : 0;
            return cst;
        };
        for (int v : vs) add(v, 0, 1);
        int cc = cost();
        for (int sweep = 0; sweep < 200 && cc > 0; sweep++) {
            bool improved = false;
            for (int v : vs) {
                int best = order[v], bestc = cc;
                add(v, order[v], -1);
                for (int pidx = 0; pidx < 6; pidx++) {
                    add(v, pidx, 1);
                    const int c2 = cost();
                    // ties move too (plateau walk), biased by the sweep so the search does not cycle
                    if (c2 < bestc || (c2 == bestc && ((rng.next() & 7) == 0))) { bestc = c2; best = pidx; }
                    add(v, pidx, -1);
                }
                add(v, best, 1);
                if (bestc < cc) improved = true;
                order[v] = best;
                cc = bestc;
            }
            (void)improved;
        }
        residual += cc;
        for (int v : vs) {
            unsigned long long pk = 0;
            for (int j = 0; j < 3; j++)
                pk |= (unsigned long long)slot_of(var_checks[3 * v + PERM[order[v]][j]], v) << (16 * j);
            out->bpack[out->posof[v]] = pk;
        }
    }
    out->residual = residual;
    return true;
}
