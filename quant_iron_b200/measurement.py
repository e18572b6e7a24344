"""MeasurementBasis / MeasurementResult (src/components/measurement.rs:15-86)."""
from __future__ import annotations


class MeasurementBasis:
    """`enum MeasurementBasis { Computational, X, Y, Custom([[Complex<f64>; 2]; 2]) }`."""
    _CODES = {"Computational": 0, "X": 1, "Y": 2, "Custom": 3}

    def __init__(self, name: str, matrix=None):
        self.name, self.matrix = name, matrix
        self.code = self._CODES[name]

    @staticmethod
    def Custom(matrix) -> "MeasurementBasis":
        return MeasurementBasis("Custom", [[complex(z) for z in row] for row in matrix])

    def flat_matrix(self):
        out = []
        for row in self.matrix:
            for z in row:
                out += [z.real, z.imag]
        return out

    def __eq__(self, other):
        return isinstance(other, MeasurementBasis) and self.name == other.name and self.matrix == other.matrix

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return f"MeasurementBasis.{self.name}"


MeasurementBasis.Computational = MeasurementBasis("Computational")
MeasurementBasis.X = MeasurementBasis("X")
MeasurementBasis.Y = MeasurementBasis("Y")


class MeasurementResult:
    """measurement.rs:15-25; unknown attributes fall through to `new_state` (Deref, 28-34)."""

    def __init__(self, basis, indices, outcomes, new_state):
        self.basis, self.indices, self.outcomes, self.new_state = basis, list(indices), list(outcomes), new_state

    def get_indices(self):
        return self.indices

    def get_basis(self):
        return self.basis

    def get_outcomes(self):
        return self.outcomes

    def get_new_state(self):
        return self.new_state

    def __getattr__(self, name):
        return getattr(self.new_state, name)
