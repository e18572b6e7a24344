//! Safe host layer over `libqiron_b200` with quant-iron's names, argument order and error variants.
//!
//! This is the code a maintainer drops into quant-iron in place of `components/state.rs`' `Vec` storage and
//! `components/operator.rs`' rayon / OpenCL dispatch (INTEGRATION.md section 3 walks through the diff).  It is committed as
//! SOURCE ONLY: the build image has no cargo / rustc, so it has never been compiled; the `-sys` crate it sits on is generated
//! from the C header and carries compile-time layout asserts, and the Python mirror (`quant_iron_b200/`), which is the same
//! layer line for line over the same C ABI, is what the parity tests drive.
//!
//! Reference citations are `file:line` in LordSaumya/quant-iron v2.0.0.
#![allow(clippy::too_many_arguments)]

use num_complex::Complex;
use quant_iron_b200_sys as sys;
use std::os::raw::c_int;
use std::ptr::{self, NonNull};

// ---- errors.rs:3-97: status codes 1..15 are the variants of `enum Error` in declaration order -------------------------
#[derive(Debug, Clone, PartialEq)]
pub enum Error {
    InvalidNumberOfMeasurements(usize),
    OverlappingControlAndTargetQubits(usize, usize),
    InvalidNumberOfQubits(usize),
    InvalidQubitIndex(usize, usize),
    StateVectorNotNormalised,
    NonUnitaryMatrix,
    InvalidNumberOfInputs(usize, usize),
    MismatchedNumberOfParameters { expected: usize, actual: usize },
    UnknownError,
    /// takes the slot of `OpenCLError(String)` (errors.rs:78): device failures
    DeviceError(String),
    GpuContextLockError,
    CircuitMacroError(String),
    InvalidInputValue(usize),
    ZeroNorm,
    InvalidPauliStringCoefficient(Complex<f64>),
    /// NULL handle / malformed record / peer failure: no reference counterpart (statuses 16, 17)
    InvalidArgument(String),
}

fn check(status: c_int) -> Result<(), Error> {
    if status == sys::QI_OK {
        return Ok(());
    }
    let mut p = [0u64; 2];
    let mut msg = [0 as std::os::raw::c_char; 256];
    unsafe { sys::qi_last_error(p.as_mut_ptr(), msg.as_mut_ptr(), msg.len()) };
    let text = unsafe { std::ffi::CStr::from_ptr(msg.as_ptr()) }.to_string_lossy().into_owned();
    Err(match status {
        1 => Error::InvalidNumberOfMeasurements(p[0] as usize),
        2 => Error::OverlappingControlAndTargetQubits(p[0] as usize, p[1] as usize),
        3 => Error::InvalidNumberOfQubits(p[0] as usize),
        4 => Error::InvalidQubitIndex(p[0] as usize, p[1] as usize),
        5 => Error::StateVectorNotNormalised,
        6 => Error::NonUnitaryMatrix,
        7 => Error::InvalidNumberOfInputs(p[0] as usize, p[1] as usize),
        8 => Error::MismatchedNumberOfParameters { expected: p[0] as usize, actual: p[1] as usize },
        10 => Error::DeviceError(text),
        11 => Error::GpuContextLockError,
        12 => Error::CircuitMacroError(text),
        13 => Error::InvalidInputValue(p[0] as usize),
        14 => Error::ZeroNorm,
        15 => Error::InvalidPauliStringCoefficient(Complex::new(f64::from_bits(p[0]), f64::from_bits(p[1]))),
        16 | 17 => Error::InvalidArgument(text),
        _ => Error::UnknownError,
    })
}

// ---- state.rs:74-81: the amplitudes live in HBM; `state_vector` becomes an accessor (documented API break) --------------
pub struct State {
    handle: NonNull<sys::qi_state>,
    pub num_qubits: usize,
}
unsafe impl Send for State {}

impl Drop for State {
    fn drop(&mut self) {
        unsafe { sys::qi_state_free(self.handle.as_ptr()) }
    }
}
impl Clone for State {
    // #[derive(Clone)] state.rs:69 -> device-to-device copy
    fn clone(&self) -> Self {
        let mut out = ptr::null_mut();
        check(unsafe { sys::qi_state_clone(self.handle.as_ptr(), &mut out) }).expect("qi_state_clone");
        State { handle: NonNull::new(out).expect("null state"), num_qubits: self.num_qubits }
    }
}

fn u32s(v: &[usize]) -> Vec<u32> {
    v.iter().map(|&q| q as u32).collect()
}

impl State {
    fn wrap(n: usize, f: impl FnOnce(*mut *mut sys::qi_state) -> c_int) -> Result<State, Error> {
        let mut out = ptr::null_mut();
        check(f(&mut out))?;
        Ok(State { handle: NonNull::new(out).ok_or(Error::UnknownError)?, num_qubits: n })
    }
    /// State::new (state.rs:100-136): length must be a power of two and the vector normalised
    pub fn new(state_vector: Vec<Complex<f64>>) -> Result<State, Error> {
        let len = state_vector.len();
        let n = if len == 0 { 0 } else { len.trailing_zeros() as usize };
        State::wrap(n, |o| unsafe { sys::qi_state_from_host(state_vector.as_ptr() as *const f64, len as u64, n as u32, 1, o) })
    }
    pub fn new_zero(n: usize) -> Result<State, Error> { State::wrap(n, |o| unsafe { sys::qi_state_new_zero(n as u32, o) }) }
    pub fn new_basis_n(n: usize, k: usize) -> Result<State, Error> { State::wrap(n, |o| unsafe { sys::qi_state_new_basis_n(n as u32, k as u64, o) }) }
    pub fn new_plus(n: usize) -> Result<State, Error> { State::wrap(n, |o| unsafe { sys::qi_state_new_plus(n as u32, o) }) }
    pub fn new_minus(n: usize) -> Result<State, Error> { State::wrap(n, |o| unsafe { sys::qi_state_new_minus(n as u32, o) }) }
    pub fn new_ghz(n: usize) -> Result<State, Error> { State::wrap(n, |o| unsafe { sys::qi_state_new_ghz(n as u32, o) }) }

    pub fn num_qubits(&self) -> usize { self.num_qubits }
    /// was the public field `state_vector` (state.rs:77): copies the amplitudes to the host
    pub fn state_vector(&self) -> Result<Vec<Complex<f64>>, Error> {
        let len = 1usize << self.num_qubits;
        let mut v = vec![Complex::new(0.0, 0.0); len];
        check(unsafe { sys::qi_state_to_host(self.handle.as_ptr(), v.as_mut_ptr() as *mut f64, len as u64) })?;
        Ok(v)
    }
    /// state.rs:448-453
    pub fn amplitude(&self, n: usize) -> Result<Complex<f64>, Error> {
        let mut z = [0.0f64; 2];
        check(unsafe { sys::qi_state_amplitude(self.handle.as_ptr(), n as u64, z.as_mut_ptr()) })?;
        Ok(Complex::new(z[0], z[1]))
    }
    /// state.rs:890-917
    pub fn inner_product(&self, other: &State) -> Result<Complex<f64>, Error> {
        let mut z = [0.0f64; 2];
        check(unsafe { sys::qi_inner_product(self.handle.as_ptr(), other.handle.as_ptr(), z.as_mut_ptr()) })?;
        Ok(Complex::new(z[0], z[1]))
    }
    /// state.rs:924-945
    pub fn normalise(&self) -> Result<State, Error> {
        let out = self.clone();
        check(unsafe { sys::qi_normalise(out.handle.as_ptr()) })?;
        Ok(out)
    }
    /// state.rs:801-836
    pub fn tensor_product(&self, other: &State) -> Result<State, Error> {
        State::wrap(self.num_qubits + other.num_qubits, |o| unsafe { sys::qi_tensor_product(self.handle.as_ptr(), other.handle.as_ptr(), o) })
    }

    /// State::operate (state.rs:1018-1031): any `Operator`, targets and controls
    pub fn operate(&self, op: &dyn Operator, targets: &[usize], controls: &[usize]) -> Result<State, Error> {
        op.apply(self, targets, controls)
    }
    fn gate(&self, kind: c_int, targets: &[usize], controls: &[usize], params: [f64; 8]) -> Result<State, Error> {
        let out = self.clone(); // `&self -> State`: device clone + in-place kernel
        out.gate_(kind, targets, controls, params)?;
        Ok(out)
    }
    /// in-place twin (a 33-qubit state cannot exist twice in 180 GB)
    fn gate_(&self, kind: c_int, targets: &[usize], controls: &[usize], params: [f64; 8]) -> Result<(), Error> {
        let c = u32s(controls);
        let g = sys::qi_gate {
            kind,
            num_targets: targets.len() as u32,
            targets: [targets.first().copied().unwrap_or(0) as u32, targets.get(1).copied().unwrap_or(0) as u32],
            num_controls: c.len() as u32,
            controls: if c.is_empty() { ptr::null() } else { c.as_ptr() },
            params,
        };
        check(unsafe { sys::qi_apply_gate(self.handle.as_ptr(), &g) }) // validation order and Error variant come from the library
    }

    // single-qubit gate methods (state.rs:1050-2330); `_multi` / controlled forms take slices as in the crate
    pub fn h(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_H, &[q], &[], [0.0; 8]) }
    pub fn x(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_X, &[q], &[], [0.0; 8]) }
    pub fn y(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_Y, &[q], &[], [0.0; 8]) }
    pub fn z(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_Z, &[q], &[], [0.0; 8]) }
    pub fn s(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_S, &[q], &[], [0.0; 8]) }
    pub fn t(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_T, &[q], &[], [0.0; 8]) }
    pub fn s_dag(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_SDG, &[q], &[], [0.0; 8]) }
    pub fn t_dag(&self, q: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_TDG, &[q], &[], [0.0; 8]) }
    pub fn p(&self, q: usize, angle: f64) -> Result<State, Error> { self.gate(sys::QI_GATE_P, &[q], &[], p1(angle)) }
    pub fn rx(&self, q: usize, angle: f64) -> Result<State, Error> { self.gate(sys::QI_GATE_RX, &[q], &[], p1(angle)) }
    pub fn ry(&self, q: usize, angle: f64) -> Result<State, Error> { self.gate(sys::QI_GATE_RY, &[q], &[], p1(angle)) }
    pub fn rz(&self, q: usize, angle: f64) -> Result<State, Error> { self.gate(sys::QI_GATE_RZ, &[q], &[], p1(angle)) }
    /// State::cnot(control, target) (state.rs:2103-2110) -- note the builder's `cnot_gate(target, control)`
    pub fn cnot(&self, control: usize, target: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_CNOT, &[target], &[control], [0.0; 8]) }
    pub fn swap(&self, a: usize, b: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_SWAP, &[a, b], &[], [0.0; 8]) }
    pub fn cswap(&self, a: usize, b: usize, controls: &[usize]) -> Result<State, Error> { self.gate(sys::QI_GATE_SWAP, &[a, b], controls, [0.0; 8]) }
    pub fn toffoli(&self, c1: usize, c2: usize, target: usize) -> Result<State, Error> { self.gate(sys::QI_GATE_TOFFOLI, &[target], &[c1, c2], [0.0; 8]) }
    pub fn unitary(&self, q: usize, m: [[Complex<f64>; 2]; 2]) -> Result<State, Error> { self.gate(sys::QI_GATE_U2, &[q], &[], mat(m)) }
    pub fn matchgate(&self, q: usize, theta: f64, phi1: f64, phi2: f64) -> Result<State, Error> {
        self.gate(sys::QI_GATE_MATCHGATE, &[q], &[], [theta, phi1, phi2, 0.0, 0.0, 0.0, 0.0, 0.0])
    }
    pub fn h_multi(&self, qs: &[usize]) -> Result<State, Error> { qs.iter().try_fold(self.clone(), |s, &q| { s.gate_(sys::QI_GATE_H, &[q], &[], [0.0; 8])?; Ok(s) }) }
    pub fn ch_multi(&self, targets: &[usize], controls: &[usize]) -> Result<State, Error> {
        targets.iter().try_fold(self.clone(), |s, &q| { s.gate_(sys::QI_GATE_H, &[q], controls, [0.0; 8])?; Ok(s) })
    }
    pub fn cp_multi(&self, targets: &[usize], controls: &[usize], angle: f64) -> Result<State, Error> {
        targets.iter().try_fold(self.clone(), |s, &q| { s.gate_(sys::QI_GATE_P, &[q], controls, p1(angle))?; Ok(s) })
    }

    /// State::measure (state.rs:525-730) with the shared-seed draw (seed, draw_index) of the oracle contract
    pub fn measure(&self, basis: MeasurementBasis, qubits: &[usize], seed: u64, draw_index: u64) -> Result<MeasurementResult, Error> {
        let out = self.clone();
        let q = u32s(qubits);
        let m = if q.is_empty() { self.num_qubits } else { q.len() };
        let mut outcomes = vec![0u8; m];
        let mut bin = 0u64;
        let (code, custom) = basis.lower();
        check(unsafe {
            sys::qi_measure(out.handle.as_ptr(), code, custom.as_ref().map_or(ptr::null(), |u| u.as_ptr()), q.as_ptr(), q.len() as u32,
                            seed, draw_index, outcomes.as_mut_ptr(), &mut bin)
        })?;
        Ok(MeasurementResult { basis, indices: qubits.to_vec(), outcomes, new_state: out })
    }
    /// State::measure_n (state.rs:750-784): n independent measurements of the same state
    pub fn measure_n(&self, basis: MeasurementBasis, qubits: &[usize], n: usize, seed: u64) -> Result<Vec<MeasurementResult>, Error> {
        if n == 0 {
            return Err(Error::InvalidNumberOfMeasurements(0));
        }
        (0..n as u64).map(|k| self.measure(basis, qubits, seed, k)).collect()
    }
    pub(crate) fn raw(&self) -> *mut sys::qi_state { self.handle.as_ptr() }
}

fn p1(a: f64) -> [f64; 8] { [a, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0] }
fn mat(m: [[Complex<f64>; 2]; 2]) -> [f64; 8] { [m[0][0].re, m[0][0].im, m[0][1].re, m[0][1].im, m[1][0].re, m[1][0].im, m[1][1].re, m[1][1].im] }

/// ChainableState (state.rs:2375-2684): gate methods on `Result<State, Error>` so that calls chain without `?`
pub trait ChainableState {
    fn h(self, q: usize) -> Result<State, Error>;
    fn x(self, q: usize) -> Result<State, Error>;
    fn cnot(self, control: usize, target: usize) -> Result<State, Error>;
    fn rx(self, q: usize, angle: f64) -> Result<State, Error>;
    fn rz(self, q: usize, angle: f64) -> Result<State, Error>;
    fn operate(self, op: &dyn Operator, targets: &[usize], controls: &[usize]) -> Result<State, Error>;
}
impl ChainableState for Result<State, Error> {
    // the chain owns its intermediate state, so every link mutates in place instead of cloning
    fn h(self, q: usize) -> Result<State, Error> { let s = self?; s.gate_(sys::QI_GATE_H, &[q], &[], [0.0; 8])?; Ok(s) }
    fn x(self, q: usize) -> Result<State, Error> { let s = self?; s.gate_(sys::QI_GATE_X, &[q], &[], [0.0; 8])?; Ok(s) }
    fn cnot(self, control: usize, target: usize) -> Result<State, Error> { let s = self?; s.gate_(sys::QI_GATE_CNOT, &[target], &[control], [0.0; 8])?; Ok(s) }
    fn rx(self, q: usize, angle: f64) -> Result<State, Error> { let s = self?; s.gate_(sys::QI_GATE_RX, &[q], &[], p1(angle))?; Ok(s) }
    fn rz(self, q: usize, angle: f64) -> Result<State, Error> { let s = self?; s.gate_(sys::QI_GATE_RZ, &[q], &[], p1(angle))?; Ok(s) }
    fn operate(self, op: &dyn Operator, targets: &[usize], controls: &[usize]) -> Result<State, Error> { op.apply(&self?, targets, controls) }
}

// ---- measurement.rs ---------------------------------------------------------------------------------------------------
#[derive(Debug, Clone, Copy, PartialEq)]
pub enum MeasurementBasis {
    Computational,
    X,
    Y,
    Custom([[Complex<f64>; 2]; 2]),
}
impl MeasurementBasis {
    fn lower(&self) -> (c_int, Option<[f64; 8]>) {
        match self {
            MeasurementBasis::Computational => (sys::QI_BASIS_COMPUTATIONAL, None),
            MeasurementBasis::X => (sys::QI_BASIS_X, None),
            MeasurementBasis::Y => (sys::QI_BASIS_Y, None),
            MeasurementBasis::Custom(u) => (sys::QI_BASIS_CUSTOM, Some(mat(*u))),
        }
    }
}
pub struct MeasurementResult {
    pub basis: MeasurementBasis,
    pub indices: Vec<usize>,
    pub outcomes: Vec<u8>,
    pub new_state: State,
}

// ---- operator.rs:150-170: the Operator trait; every built-in operator is one gate record ---------------------------------
pub trait Operator {
    fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error>;
    fn base_qubits(&self) -> usize;
    /// the record `Circuit::execute` batches (None: a user-defined operator, applied through `apply`)
    fn record(&self) -> Option<(c_int, [f64; 8])> { None }
}
macro_rules! fixed_gate {
    ($name:ident, $kind:expr, $base:expr) => {
        #[derive(Debug, Clone, Copy)]
        pub struct $name;
        impl Operator for $name {
            fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error> { state.gate($kind, targets, controls, [0.0; 8]) }
            fn base_qubits(&self) -> usize { $base }
            fn record(&self) -> Option<(c_int, [f64; 8])> { Some(($kind, [0.0; 8])) }
        }
    };
}
fixed_gate!(Hadamard, sys::QI_GATE_H, 1); // operator.rs:303-424
fixed_gate!(Identity, sys::QI_GATE_I, 1); // operator.rs:1112-1123
fixed_gate!(PhaseS, sys::QI_GATE_S, 1);
fixed_gate!(PhaseT, sys::QI_GATE_T, 1);
fixed_gate!(PhaseSdag, sys::QI_GATE_SDG, 1);
fixed_gate!(PhaseTdag, sys::QI_GATE_TDG, 1);
fixed_gate!(CNOT, sys::QI_GATE_CNOT, 2); // operator.rs:667-685 (targets = [target], controls = [control])
fixed_gate!(SWAP, sys::QI_GATE_SWAP, 2); // operator.rs:731-820
fixed_gate!(Toffoli, sys::QI_GATE_TOFFOLI, 3); // operator.rs:1055-1075

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Pauli { I = 0, X = 1, Y = 2, Z = 3 }
impl Operator for Pauli {
    // operator.rs:474-606
    fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error> {
        let (kind, _) = self.record().unwrap();
        state.gate(kind, targets, controls, [0.0; 8])
    }
    fn base_qubits(&self) -> usize { 1 }
    fn record(&self) -> Option<(c_int, [f64; 8])> {
        Some((match self { Pauli::I => sys::QI_GATE_I, Pauli::X => sys::QI_GATE_X, Pauli::Y => sys::QI_GATE_Y, Pauli::Z => sys::QI_GATE_Z }, [0.0; 8]))
    }
}
macro_rules! angle_gate {
    ($name:ident, $kind:expr) => {
        #[derive(Debug, Clone, Copy)]
        pub struct $name { pub angle: f64 }
        impl $name { pub fn new(angle: f64) -> Self { $name { angle } } }
        impl Operator for $name {
            fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error> { state.gate($kind, targets, controls, p1(self.angle)) }
            fn base_qubits(&self) -> usize { 1 }
            fn record(&self) -> Option<(c_int, [f64; 8])> { Some(($kind, p1(self.angle))) }
        }
    };
}
angle_gate!(PhaseShift, sys::QI_GATE_P); // operator.rs:1565-1624
angle_gate!(RotateX, sys::QI_GATE_RX); // operator.rs:1674-1767
angle_gate!(RotateY, sys::QI_GATE_RY); // operator.rs:1817-1908
angle_gate!(RotateZ, sys::QI_GATE_RZ); // operator.rs:1958-2035

/// Unitary2 (operator.rs:2090-2266): `new` checks unitarity with the library's tolerance (qi_unitary2_check)
#[derive(Debug, Clone, Copy)]
pub struct Unitary2 { pub matrix: [[Complex<f64>; 2]; 2] }
impl Unitary2 {
    pub fn new(matrix: [[Complex<f64>; 2]; 2]) -> Result<Self, Error> {
        check(unsafe { sys::qi_unitary2_check(mat(matrix).as_ptr()) })?;
        Ok(Unitary2 { matrix })
    }
}
impl Operator for Unitary2 {
    fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error> { state.gate(sys::QI_GATE_U2, targets, controls, mat(self.matrix)) }
    fn base_qubits(&self) -> usize { 1 }
    fn record(&self) -> Option<(c_int, [f64; 8])> { Some((sys::QI_GATE_U2, mat(self.matrix))) }
}
/// Matchgate (operator.rs:893-1014) on (target, target + 1)
#[derive(Debug, Clone, Copy)]
pub struct Matchgate { pub theta: f64, pub phi1: f64, pub phi2: f64 }
impl Operator for Matchgate {
    fn apply(&self, state: &State, targets: &[usize], controls: &[usize]) -> Result<State, Error> { state.gate(sys::QI_GATE_MATCHGATE, targets, controls, self.record().unwrap().1) }
    fn base_qubits(&self) -> usize { 2 }
    fn record(&self) -> Option<(c_int, [f64; 8])> { Some((sys::QI_GATE_MATCHGATE, [self.theta, self.phi1, self.phi2, 0.0, 0.0, 0.0, 0.0, 0.0])) }
}

// ---- pauli_string.rs ----------------------------------------------------------------------------------------------------
#[derive(Debug, Clone, PartialEq)]
pub struct PauliString { pub ops: Vec<(usize, Pauli)>, pub coefficient: Complex<f64> }
impl PauliString {
    pub fn new(coefficient: Complex<f64>) -> Self { PauliString { ops: Vec::new(), coefficient } }
    pub fn with_op(mut self, qubit: usize, op: Pauli) -> Self { self.add_op(qubit, op); self }
    /// pauli_string.rs:75-87: a later op on the same qubit replaces the earlier one (HashMap insert)
    pub fn add_op(&mut self, qubit: usize, op: Pauli) {
        self.ops.retain(|&(q, _)| q != qubit);
        self.ops.push((qubit, op));
    }
    fn with_term<R>(&self, f: impl FnOnce(&sys::qi_pauli_term) -> R) -> R {
        let live: Vec<&(usize, Pauli)> = self.ops.iter().filter(|(_, p)| *p != Pauli::I).collect();
        let q: Vec<u32> = live.iter().map(|(q, _)| *q as u32).collect();
        let p: Vec<u8> = live.iter().map(|(_, p)| *p as u8).collect();
        f(&sys::qi_pauli_term { num_ops: q.len() as u32, qubits: q.as_ptr(), paulis: p.as_ptr(), coefficient: [self.coefficient.re, self.coefficient.im] })
    }
    /// pauli_string.rs:139-151
    pub fn apply(&self, state: &State) -> Result<State, Error> {
        let out = state.clone();
        self.with_term(|t| check(unsafe { sys::qi_apply_pauli_string(out.raw(), t, 1) }))?;
        Ok(out)
    }
    /// pauli_string.rs:237-262: exp(coefficient * factor * P)
    pub fn apply_exp_factor(&self, state: &State, factor: Complex<f64>) -> Result<State, Error> {
        let out = state.clone();
        self.with_term(|t| check(unsafe { sys::qi_apply_pauli_exp(out.raw(), t, [factor.re, factor.im].as_ptr()) }))?;
        Ok(out)
    }
    pub fn apply_exp(&self, state: &State) -> Result<State, Error> { self.apply_exp_factor(state, Complex::new(1.0, 0.0)) }
}
#[derive(Debug, Clone, PartialEq)]
pub struct SumOp { pub terms: Vec<PauliString> }
impl SumOp {
    pub fn new(terms: Vec<PauliString>) -> Self { SumOp { terms } }
    fn with_terms<R>(&self, f: impl FnOnce(&[sys::qi_pauli_term]) -> R) -> R {
        // the qubit / pauli arrays must outlive the records that point into them
        let store: Vec<(Vec<u32>, Vec<u8>)> = self.terms.iter().map(|t| {
            let live: Vec<&(usize, Pauli)> = t.ops.iter().filter(|(_, p)| *p != Pauli::I).collect();
            (live.iter().map(|(q, _)| *q as u32).collect(), live.iter().map(|(_, p)| *p as u8).collect())
        }).collect();
        let recs: Vec<sys::qi_pauli_term> = self.terms.iter().zip(&store).map(|(t, (q, p))| sys::qi_pauli_term {
            num_ops: q.len() as u32, qubits: q.as_ptr(), paulis: p.as_ptr(), coefficient: [t.coefficient.re, t.coefficient.im] }).collect();
        f(&recs)
    }
    /// pauli_string.rs:485-507: one batched call (terms that fit a register window share one read pass)
    pub fn expectation_value(&self, state: &State) -> Result<Complex<f64>, Error> {
        let mut z = [0.0f64; 2];
        self.with_terms(|r| check(unsafe { sys::qi_expect_pauli_sum(state.raw(), r.as_ptr(), r.len() as u64, z.as_mut_ptr()) }))?;
        Ok(Complex::new(z[0], z[1]))
    }
    /// pauli_string.rs:440-470
    pub fn apply(&self, state: &State) -> Result<State, Error> {
        State::wrap(state.num_qubits, |o| self.with_terms(|r| unsafe { sys::qi_apply_pauli_sum(state.raw(), r.as_ptr(), r.len() as u64, o) }))
    }
    /// time_evolution.rs:140-167: the step loops run inside the library as one fused sequence
    pub fn trotter_evolve_state(&self, state: &State, dt: f64, steps: usize, second_order: bool) -> Result<State, Error> {
        let out = state.clone();
        self.with_terms(|r| check(unsafe { sys::qi_trotter_evolve(out.raw(), r.as_ptr(), r.len() as u64, dt, steps as u64, if second_order { 2 } else { 1 }) }))?;
        Ok(out)
    }
}

// ---- gate.rs / circuit.rs -----------------------------------------------------------------------------------------------
pub enum Gate {
    Operator(Box<dyn Operator>, Vec<usize>, Vec<usize>), // gate.rs:23
    Measurement(MeasurementBasis, Vec<usize>),           // gate.rs:29
    PauliString(PauliString),
    PauliTimeEvolution(PauliString, f64),                // gate.rs:116-118: exp(-i t P)
}
pub struct Circuit { pub gates: Vec<Gate>, pub num_qubits: usize }
impl Circuit {
    pub fn new(num_qubits: usize) -> Self { Circuit { gates: Vec::new(), num_qubits } }
    pub fn add_gate(&mut self, gate: Gate) { self.gates.push(gate) }
    /// Circuit::execute (circuit.rs:160-172): ONE library call per run of operator gates -- the scheduler fuses the run into
    /// CTA-tile passes (modules) -- instead of one state sweep per gate
    pub fn execute(&self, initial_state: &State) -> Result<State, Error> {
        if initial_state.num_qubits() != self.num_qubits {
            return Err(Error::InvalidNumberOfQubits(initial_state.num_qubits()));
        }
        let out = initial_state.clone();
        let mut run: Vec<sys::qi_gate> = Vec::new();
        let mut ctrl_store: Vec<Vec<u32>> = Vec::new(); // control lists the records point into
        let flush = |run: &mut Vec<sys::qi_gate>, store: &mut Vec<Vec<u32>>| -> Result<(), Error> {
            if !run.is_empty() {
                check(unsafe { sys::qi_apply_circuit(out.raw(), run.as_ptr(), run.len() as u64) })?;
            }
            run.clear();
            store.clear();
            Ok(())
        };
        let mut draw = 0u64;
        for gate in &self.gates {
            match gate {
                Gate::Operator(op, targets, controls) => match op.record() {
                    Some((kind, params)) => {
                        ctrl_store.push(u32s(controls));
                        let c = ctrl_store.last().unwrap();
                        run.push(sys::qi_gate { kind, num_targets: targets.len() as u32,
                                                targets: [targets.first().copied().unwrap_or(0) as u32, targets.get(1).copied().unwrap_or(0) as u32],
                                                num_controls: c.len() as u32, controls: if c.is_empty() { ptr::null() } else { c.as_ptr() }, params });
                    }
                    None => {
                        // user-defined operator (circuit.rs:1215-1224): composes the same primitives through `apply`
                        flush(&mut run, &mut ctrl_store)?;
                        let next = op.apply(&out, targets, controls)?;
                        check(unsafe { sys::qi_scale(out.raw(), [0.0, 0.0].as_ptr()) })?;
                        check(unsafe { sys::qi_add(out.raw(), next.raw()) })?;
                    }
                },
                Gate::Measurement(basis, qubits) => {
                    flush(&mut run, &mut ctrl_store)?;
                    let q = u32s(qubits);
                    let (code, custom) = basis.lower();
                    let mut outcomes = vec![0u8; if q.is_empty() { self.num_qubits } else { q.len() }];
                    let mut bin = 0u64;
                    check(unsafe { sys::qi_measure(out.raw(), code, custom.as_ref().map_or(ptr::null(), |u| u.as_ptr()), q.as_ptr(), q.len() as u32,
                                                   0, draw, outcomes.as_mut_ptr(), &mut bin) })?;
                    draw += 1;
                }
                Gate::PauliString(ps) => {
                    flush(&mut run, &mut ctrl_store)?;
                    ps.with_term(|t| check(unsafe { sys::qi_apply_pauli_string(out.raw(), t, 1) }))?;
                }
                Gate::PauliTimeEvolution(ps, time) => {
                    flush(&mut run, &mut ctrl_store)?;
                    ps.with_term(|t| check(unsafe { sys::qi_apply_pauli_exp(out.raw(), t, [0.0, -*time].as_ptr()) }))?;
                }
            }
        }
        flush(&mut run, &mut ctrl_store)?;
        Ok(out)
    }
}
/// CircuitBuilder (circuit.rs:300-1300): the adders the benchmarks and the QFT need; `cnot_gate(target, control)` as in circuit.rs:1071-1075
pub struct CircuitBuilder { circuit: Circuit }
impl CircuitBuilder {
    pub fn new(num_qubits: usize) -> Self { CircuitBuilder { circuit: Circuit::new(num_qubits) } }
    fn op(&mut self, op: Box<dyn Operator>, targets: &[usize], controls: &[usize]) -> &mut Self {
        self.circuit.add_gate(Gate::Operator(op, targets.to_vec(), controls.to_vec()));
        self
    }
    pub fn h_gate(&mut self, q: usize) -> &mut Self { self.op(Box::new(Hadamard), &[q], &[]) }
    pub fn x_gate(&mut self, q: usize) -> &mut Self { self.op(Box::new(Pauli::X), &[q], &[]) }
    pub fn rx_gate(&mut self, q: usize, angle: f64) -> &mut Self { self.op(Box::new(RotateX::new(angle)), &[q], &[]) }
    pub fn ry_gate(&mut self, q: usize, angle: f64) -> &mut Self { self.op(Box::new(RotateY::new(angle)), &[q], &[]) }
    pub fn rz_gate(&mut self, q: usize, angle: f64) -> &mut Self { self.op(Box::new(RotateZ::new(angle)), &[q], &[]) }
    pub fn cnot_gate(&mut self, target: usize, control: usize) -> &mut Self { self.op(Box::new(CNOT), &[target], &[control]) }
    pub fn cp_gates(&mut self, targets: &[usize], controls: &[usize], angle: f64) -> &mut Self {
        for &t in targets { self.op(Box::new(PhaseShift::new(angle)), &[t], controls); }
        self
    }
    pub fn swap_gate(&mut self, a: usize, b: usize) -> &mut Self { self.op(Box::new(SWAP), &[a, b], &[]) }
    pub fn toffoli_gate(&mut self, c1: usize, c2: usize, target: usize) -> &mut Self { self.op(Box::new(Toffoli), &[target], &[c1, c2]) }
    pub fn add_operator_gate(&mut self, op: Box<dyn Operator>, targets: &[usize], controls: &[usize]) -> &mut Self { self.op(op, targets, controls) }
    pub fn measure_gate(&mut self, basis: MeasurementBasis, qubits: &[usize]) -> &mut Self { self.circuit.add_gate(Gate::Measurement(basis, qubits.to_vec())); self }
    pub fn pauli_time_evolution_gate(&mut self, ps: PauliString, time: f64) -> &mut Self { self.circuit.add_gate(Gate::PauliTimeEvolution(ps, time)); self }
    /// Subroutine::qft (subroutine.rs:90-112)
    pub fn qft(&mut self, qubits: &[usize]) -> &mut Self {
        let n = qubits.len();
        for i in 0..n {
            self.h_gate(qubits[i]);
            for k in 1..(n - i) {
                self.cp_gates(&[qubits[i]], &[qubits[i + k]], std::f64::consts::PI / (1u64 << k) as f64);
            }
        }
        for i in 0..n / 2 {
            self.swap_gate(qubits[i], qubits[n - 1 - i]);
        }
        self
    }
    /// build_final (circuit.rs:352-355): hands the gate list over and leaves the builder empty and reusable.  (`build`,
    /// circuit.rs:340-343, clones the list instead; `Box<dyn Operator>` would need the crate's `box_clone` for that.)
    pub fn build_final(&mut self) -> Result<Circuit, Error> {
        let n = self.circuit.num_qubits;
        Ok(std::mem::replace(&mut self.circuit, Circuit::new(n)))
    }
}
