// Host-only logic of the C++ facade (include/quant_iron_b200.hpp): model builders, builder validation, the
// unitarity check -- everything that runs without a device, so this binary is part of the CPU test tier.
// Facts restated from src/tests/{ising,heisenberg,circuit,operator}_tests.rs (cited).
#include <cstdio>
#include <cstdlib>

#include "quant_iron_b200.hpp"

using namespace quant_iron;
static int failures = 0;
#define EXPECT(cond)                                                                 \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

template <class F>
static bool throws(const char* variant, uint64_t p0, uint64_t p1, F f) {
    try { f(); } catch (const Error& e) { return e.variant == variant && e.payload[0] == p0 && e.payload[1] == p1; }
    return false;
}

static bool has_term(const SumOp& s, std::vector<std::pair<size_t, Pauli>> ops, double coeff) {
    for (auto& t : s.terms) {
        if (t.len() != ops.size() || std::abs(t.coefficient() - cplx(coeff, 0.0)) > 1e-15) continue;
        bool all = true;
        for (auto& o : ops) { auto it = t.ops().find(o.first); all = all && it != t.ops().end() && it->second == o.second; }
        if (all) return true;
    }
    return false;
}

int main() {
    // heisenberg_tests.rs:13-31 / heisenberg.rs:46-49: 4 sites -> 16 terms, per site XX, YY, ZZ, Z; field +mu*h/2
    SumOp h1 = heisenberg_1d(4, 1.0, 2.0, 3.0, 4.0, 5.0);
    EXPECT(h1.num_terms() == 16);
    EXPECT(has_term(h1, {{0, Pauli::X}, {1, Pauli::X}}, -0.5) && has_term(h1, {{3, Pauli::Y}, {0, Pauli::Y}}, -1.0));
    EXPECT(has_term(h1, {{2, Pauli::Z}}, 10.0));
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { heisenberg_1d(1, 1, 2, 3, 4, 5); }));
    // heisenberg_tests.rs:106-227: 3x3 -> 63 terms, field -0.5*h*mu, vertical and horizontal periodic bonds
    SumOp h2 = heisenberg_2d(3, 3, 1.0, 2.0, 3.0, 4.0, 5.0);
    EXPECT(h2.num_terms() == 63);
    EXPECT(has_term(h2, {{0, Pauli::Z}}, -10.0) && has_term(h2, {{7, Pauli::Z}, {1, Pauli::Z}}, -1.5) && has_term(h2, {{8, Pauli::X}, {6, Pauli::X}}, -0.5));
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { heisenberg_2d(1, 2, 1, 2, 3, 4, 5); }));
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { heisenberg_2d(2, 1, 1, 2, 3, 4, 5); }));
    EXPECT(heisenberg_2d(2, 2, 0, 0, 0, 0, 5).num_terms() == 0);
    // ising_tests.rs:13-101
    SumOp i1 = ising_1d({1.0, 2.0, 3.0}, {0.5, 1.0, 1.5}, 0.1);
    EXPECT(i1.num_terms() == 6);
    EXPECT(has_term(i1, {{0, Pauli::Z}}, -0.1) && has_term(i1, {{2, Pauli::Z}}, -0.1 * 3.0) && has_term(i1, {{2, Pauli::Z}, {0, Pauli::Z}}, -1.5));
    EXPECT(i1.terms[0].len() == 2 && i1.terms[1].len() == 1);          // coupling first, then field (ising.rs:49-66)
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { ising_1d({1.0}, {0.5}, 0.1); }));
    EXPECT(ising_1d_uniform(3, 1.0, 2.0, 0.1).num_terms() == 6 && ising_1d_uniform(4, 0.0, 0.0, 0.1).num_terms() == 0);
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { ising_1d_uniform(1, 1.0, 2.0, 0.1); }));
    // ising_tests.rs:103-316
    SumOp i2 = ising_2d({{1, 2, 3}, {4, 5, 6}, {7, 8, 9}}, {{0.5, 1.5, 2.5}, {3.5, 4.5, 5.5}, {6.5, 7.5, 8.5}},
                        {{1, 2, 3}, {4, 5, 6}, {7, 8, 9}}, 0.1);
    EXPECT(i2.num_terms() == 27);
    EXPECT(has_term(i2, {{0, Pauli::Z}, {3, Pauli::Z}}, -0.5) && has_term(i2, {{0, Pauli::Z}, {1, Pauli::Z}}, -1.0));
    EXPECT(has_term(i2, {{8, Pauli::Z}, {2, Pauli::Z}}, -8.5) && has_term(i2, {{8, Pauli::Z}, {6, Pauli::Z}}, -9.0));
    EXPECT(ising_2d_uniform(3, 3, 1.0, 2.0, 0.1).num_terms() == 27);
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { ising_2d_uniform(1, 1, 1.0, 2.0, 0.1); }));
    EXPECT(throws("InvalidNumberOfInputs", 1, 2, [] { ising_2d_uniform(2, 1, 1.0, 2.0, 0.1); }));
    // circuit_tests.rs:197-207: an out-of-range qubit is reported when the circuit is built
    EXPECT(throws("InvalidQubitIndex", 3, 2, [] { CircuitBuilder(2).h_gate(0).cnot_gate(0, 3).build(); }));
    EXPECT(CircuitBuilder(2).h_gate(0).cnot_gate(0, 1).build().gates.size() == 2);
    // subroutine.rs:90-112: qft over n qubits = n H + n(n-1)/2 CP + n/2 SWAP
    std::vector<size_t> qs = {0, 1, 2, 3, 4};
    EXPECT(Subroutine::qft(qs, 5).gates.size() == 5 + 10 + 2 && Subroutine::iqft(qs, 5).gates.size() == 17);
    // operator_tests.rs:1790-1805: Unitary2::new rejects a non-unitary matrix (host-side check, no device)
    bool rejected = false;
    const cplx bad[2][2] = {{cplx(1, 0), cplx(1, 0)}, {cplx(0, 0), cplx(1, 0)}};
    try { Unitary2::make(bad); } catch (const Error& e) { rejected = e.variant == "NonUnitaryMatrix"; }
    const cplx good[2][2] = {{cplx(0, 0), cplx(1, 0)}, {cplx(1, 0), cplx(0, 0)}};
    EXPECT(Unitary2::make(good).kind() == QI_GATE_U2);
    EXPECT(rejected);
    std::printf(failures ? "C++ facade host logic: %d FAILURES\n" : "C++ facade host logic: ALL PASS\n", failures);
    return failures ? 1 : 0;
}
