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

    def __eq__(self, other):
        return isinstance(other, Error) and (self.variant, self.payload) == (other.variant, other.payload)

    def __hash__(self):
        return hash((self.variant, self.payload))
