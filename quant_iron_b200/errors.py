"""quant-iron's `enum Error` (errors.rs:3-97) as a Python exception.

`variant` is the Rust variant name and `payload` its fields, so host code can match on them the way
the reference's tests match on `Err(Error::InvalidQubitIndex(2, 2))`.
"""


class Error(Exception):
    def __init__(self, variant: str, *payload):
        super().__init__(f"{variant}{tuple(payload)}")
        self.variant = variant
        self.payload = tuple(payload)
        self.message = ""

    # thiserror Display strings (errors.rs:10-98)
    _DISPLAY = {
        "InvalidNumberOfMeasurements": "Invalid number of measurements: {0}",
        "OverlappingControlAndTargetQubits": "Control qubit index {0} overlaps with target qubit index {1}",
        "InvalidNumberOfQubits": "Invalid number of qubits: {0}",
        "InvalidQubitIndex": "Invalid qubit index: {0} for {1} qubits",
        "StateVectorNotNormalised": "State vector is not normalised",
        "NonUnitaryMatrix": "Non-unitary matrix",
        "InvalidNumberOfInputs": "Unexpected number of inputs: expected {1}, got {0}",
        "MismatchedNumberOfParameters": "Mismatched number of parameters: expected {0}, got {1}",
        "UnknownError": "An unknown error occurred",
        "CircuitMacroError": "Failed to create circuit from macro: {0}",
        "InvalidInputValue": "Invalid input value for operation: {0}",
        "ZeroNorm": "The state cannot be normalised because it has zero norm.",
        "InvalidPauliStringCoefficient": "Invalid Pauli String coefficient: {0}",
    }

    def to_string(self) -> str:
        fmt = self._DISPLAY.get(self.variant)
        try:
            return fmt.format(*self.payload) if fmt else str(self)
        except IndexError:
            return str(self)

    def __eq__(self, other):
        return isinstance(other, Error) and (self.variant, self.payload) == (other.variant, other.payload)

    def __hash__(self):
        return hash((self.variant, self.payload))


class CompilerError(Exception):
    """`enum CompilerError { IOError, UnsupportedOperator, InvalidOperands }` (errors.rs:99-108)."""
    _DISPLAY = {"IOError": "I/O error: {0}", "UnsupportedOperator": "An unsupported operation was encountered: {0}",
                "InvalidOperands": "Invalid operands ({0}) for operator {1}"}

    def __init__(self, variant: str, *payload):
        super().__init__(f"{variant}{tuple(payload)}")
        self.variant = variant
        self.payload = tuple(payload)

    def to_string(self) -> str:
        return self._DISPLAY[self.variant].format(*self.payload)

    def __eq__(self, other):
        return isinstance(other, CompilerError) and (self.variant, self.payload) == (other.variant, other.payload)

    def __hash__(self):
        return hash((self.variant, self.payload))
