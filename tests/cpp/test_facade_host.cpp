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
    // the builder's full adder surface (circuit.rs:378-1224, 1737), counted as the reference's macro tests count it
    // (macros_tests.rs:47-74, 134-211, 231-247, 267-305): one gate per target, validation in build / build_final
    {
        using Q = std::vector<size_t>;
        const cplx xm[2][2] = {{0.0, 1.0}, {1.0, 0.0}};
        CircuitBuilder b(3);
        b.h_gate(0).h_gates({0, 1}).x_gate(1).x_gates({1, 2}).y_gate(2).y_gates({0, 2}).z_gate(0).z_gates({0, 1}).s_gate(1).s_gates({1, 2})
            .t_gate(2).t_gates({0, 2}).id_gate(0).id_gates({0, 1}).sdag_gate(1).sdag_gates({1, 2}).tdag_gate(2).tdag_gates({0, 2});
        EXPECT(b.build().gates.size() == 27);
        EXPECT(throws("InvalidQubitIndex", 3, 3, [&] { CircuitBuilder(3).tdag_gates({0, 3}).build_final(); }));
        CircuitBuilder c4(4);
        auto four = [&](auto add) { add(Q{0}, Q{1}); add(Q{0, 1}, Q{2}); add(Q{0}, Q{1, 2}); add(Q{0, 1}, Q{2, 3}); };
        four([&](Q t, Q c) { c4.ch_gates(t, c); });
        four([&](Q t, Q c) { c4.cx_gates(t, c); });
        four([&](Q t, Q c) { c4.cy_gates(t, c); });
        four([&](Q t, Q c) { c4.cz_gates(t, c); });
        four([&](Q t, Q c) { c4.cs_gates(t, c); });
        four([&](Q t, Q c) { c4.ct_gates(t, c); });
        four([&](Q t, Q c) { c4.csdag_gates(t, c); });
        four([&](Q t, Q c) { c4.ctdag_gates(t, c); });
        four([&](Q t, Q c) { c4.crx_gates(t, c, M_PI / 4); });
        four([&](Q t, Q c) { c4.cry_gates(t, c, M_PI / 4); });
        four([&](Q t, Q c) { c4.cry_phase_gates(t, c, M_PI / 4, M_PI / 2); });
        c4.cmatchgate(0, {1}, M_PI / 4, M_PI / 2, M_PI / 3).cmatchgate(0, {1, 2}, M_PI / 4, M_PI / 2, M_PI / 3);
        four([&](Q t, Q c) { c4.crz_gates(t, c, M_PI / 4); });
        four([&](Q t, Q c) { c4.cp_gates(t, c, M_PI / 4); });
        EXPECT(c4.build().gates.size() == 6 * 13 + 2);
        EXPECT(throws("InvalidQubitIndex", 6, 4, [&] { CircuitBuilder(4).cx_gates({0, 1}, {2, 6}).build(); }));
        CircuitBuilder u(4);
        u.unitary_gate(0, xm).unitary_gates({0, 1}, xm).cunitary_gates({0}, {1}, xm).cunitary_gates({0, 1}, {2}, xm).cunitary_gates({0}, {1, 2}, xm)
            .cunitary_gates({0, 1}, {2, 3}, xm);
        EXPECT(u.build().gates.size() == 9);
        const cplx bad[2][2] = {{1.0, 1.0}, {0.0, 1.0}};
        EXPECT(throws("NonUnitaryMatrix", 0, 0, [&] { CircuitBuilder(2).unitary_gate(0, bad); }));
        CircuitBuilder m(3);
        m.measure_gate(MeasurementBasis::X, {0}).measure_gate(MeasurementBasis::X, {1, 2}).measure_gate(MeasurementBasis::Y, {0})
            .measure_gate(MeasurementBasis::Computational, {1, 2}).measure_gate(Basis::custom(xm), {0}).measure_gate(Basis::custom(xm), {1, 2});
        Circuit mc = m.build();
        EXPECT(mc.gates.size() == 6 && mc.gates[4].kind == Gate::Kind::Measurement && mc.gates[4].basis.kind == MeasurementBasis::Custom);
        CircuitBuilder a(3);
        a.rx_gate(0, 1.0).ry_gate(1, 1.0).rz_gate(2, 1.0).p_gate(0, 1.0).ry_phase_gate(1, 1.0, 0.5).matchgate(0, 1.0, 0.5, 0.25).rx_gates({0, 1}, 1.0)
            .ry_gates({1, 2}, 1.0).rz_gates({0, 2}, 1.0).ry_phase_gates({0, 1}, 1.0, 0.5).p_gates({0, 1, 2}, 1.0);
        EXPECT(a.build().gates.size() == 4 + 2 + 2 + 2 + 3 + 4);
        CircuitBuilder t5(5);
        t5.cnot_gate(0, 1).swap_gate(1, 2).cswap_gate(0, 1, {2}).cswap_gate(0, 1, {2, 3}).toffoli_gate(0, 1, 2);
        EXPECT(t5.build().gates.size() == 5);
        EXPECT(throws("InvalidQubitIndex", 6, 5, [&] { CircuitBuilder(5).cswap_gate(0, 1, {2, 6}).build(); }));
        // Pauli gates carry their targets (gate.rs:38-52); build_final empties the builder (circuit.rs:352-358)
        PauliString ps = PauliString(cplx(0.5, 0.0)).with_op(0, Pauli::X).with_op(2, Pauli::Z);
        CircuitBuilder pb(3);
        pb.pauli_string_gate(ps).pauli_time_evolution_gate(ps, 0.1);
        Circuit pc = pb.build_final();
        EXPECT(pc.gates.size() == 2 && pc.gates[0].kind == Gate::Kind::PauliString && pc.gates[1].kind == Gate::Kind::PauliTimeEvolution);
        EXPECT(pc.gates[1].targets.size() == 2 && pc.gates[1].time == 0.1 && pb.build().gates.empty());
        EXPECT(throws("InvalidQubitIndex", 5, 3, [&] { CircuitBuilder(3).pauli_string_gate(PauliString(cplx(1.0, 0.0)).with_op(5, Pauli::Y)).build(); }));
        // ry_phase / ry_phase_dagger are inverses of each other (operator.rs:2140-2192): U^dagger = dagger form
        Unitary2 f = Unitary2::from_ry_phase(0.7, 0.3), d = Unitary2::from_ry_phase_dagger(0.7, 0.3);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) EXPECT(std::abs(std::conj(f.m[j][i]) - d.m[i][j]) < 1e-15);
    }
    // Parameter / parametric gates (components/parametric/*.rs; parametric_tests.rs:9-60, 238-283): a clone shares the cell,
    // a deep clone does not; a parametric gate resolves with the values of the moment it is resolved
    {
        Parameter<1> th({0.25});
        Parameter<1> shared = th.clone(), copy = th.deep_clone();
        th.set({0.75});
        EXPECT(shared.get()[0] == 0.75 && copy.get()[0] == 0.25);
        Circuit pc = CircuitBuilder(3).parametric_rx_gate(0, th).parametric_cry_phase_gates({1, 2}, {0}, {Parameter<2>({0.1, 0.2}), Parameter<2>({0.3, 0.4})})
                         .parametric_cmatchgate(0, {2}, Parameter<3>({0.5, 0.6, 0.7})).h_gate(1).build();
        EXPECT(pc.gates.size() == 5 && pc.gates[0].kind == Gate::Kind::Parametric);
        Circuit cc = pc.to_concrete_circuit();
        EXPECT(cc.gates.size() == 5 && cc.gates[0].kind == Gate::Kind::Operator && cc.gates[0].op->kind() == QI_GATE_RX && cc.gates[0].op->params()[0] == 0.75);
        EXPECT(cc.gates[3].op->kind() == QI_GATE_MATCHGATE && cc.gates[3].controls == std::vector<size_t>{2} && cc.gates[3].op->params()[2] == 0.7);
        th.set({1.5});
        EXPECT(pc.to_concrete_circuit().gates[0].op->params()[0] == 1.5 && cc.gates[0].op->params()[0] == 0.75);
        EXPECT(throws("MismatchedNumberOfParameters", 2, 1, [&] { CircuitBuilder(3).parametric_rx_gates({0, 1}, {th}); }));
        // a PauliString gate becomes its factors' Pauli gates (circuit.rs:205-217, pauli_string.rs:118-122)
        Circuit ps = CircuitBuilder(3).pauli_string_gate(PauliString(cplx(2.0, 0.0)).with_op(0, Pauli::X).with_op(2, Pauli::Y)).build().to_concrete_circuit();
        EXPECT(ps.gates.size() == 2 && ps.gates[0].op->kind() == QI_GATE_X && ps.gates[1].op->kind() == QI_GATE_Y && ps.gates[1].targets[0] == 2);
    }
    std::printf(failures ? "C++ facade host logic: %d FAILURES\n" : "C++ facade host logic: ALL PASS\n", failures);
    return failures ? 1 : 0;
}
