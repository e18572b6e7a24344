"""ChainableState (state.rs:2375-2684): gate methods on `Result<State, Error>`.

The reference lets calls chain without `?` -- `State::new_zero(2).h(0).cnot(0, 1)` -- by implementing every gate method on
`Result<State, Error>` as `self.and_then(|state| state.method(..))` (state.rs:2611-2620): the first error short-circuits the
rest of the chain and is what the chain evaluates to.  The Python mirror raises instead of returning `Result`, so a chain of
plain `State` methods already stops at the first error; `ChainableState` is the Result-shaped form for callers that want the
reference's semantics literally: every State method is forwarded while the chain is Ok, skipped once it is Err.
"""
from __future__ import annotations

from .errors import Error


class ChainableState:
    """`Result<State, Error>` with the State methods on it.  `chain(State.new_zero(2)).h(0).cnot(0, 1).unwrap()`."""

    __slots__ = ("_state", "_error")

    def __init__(self, state=None, error: Error | None = None):
        self._state, self._error = state, error

    # Result<_, _> surface
    def is_ok(self) -> bool:
        return self._error is None

    def is_err(self) -> bool:
        return self._error is not None

    def unwrap(self):
        if self._error is not None:
            raise self._error
        return self._state

    def unwrap_err(self) -> Error:
        if self._error is None:
            raise ValueError("called unwrap_err on an Ok value")
        return self._error

    def ok(self):
        return self._state if self._error is None else None

    def err(self):
        return self._error

    def and_then(self, fn) -> "ChainableState":
        """`Result::and_then`: fn(State) -> State | ChainableState, errors raised by it become the chain's value."""
        if self._error is not None:
            return self
        try:
            out = fn(self._state)
        except Exception as ex:  # noqa: BLE001  (any implementation's Error type: the oracle has its own class)
            if type(ex).__name__ != "Error":
                raise
            return ChainableState(None, ex)
        return out if isinstance(out, ChainableState) else ChainableState(out)

    def __getattr__(self, name):
        # every gate / operate / measure method of State (state.rs:2623-2684 lists them one by one)
        if name.startswith("_"):
            raise AttributeError(name)

        def call(*args, **kwargs):
            return self.and_then(lambda st: getattr(st, name)(*args, **kwargs))
        return call

    def __repr__(self):
        return f"Ok({self._state!r})" if self._error is None else f"Err({self._error!r})"


def chain(state_or_result) -> ChainableState:
    """Start a chain from a State (Ok) or from an existing ChainableState."""
    return state_or_result if isinstance(state_or_result, ChainableState) else ChainableState(state_or_result)
