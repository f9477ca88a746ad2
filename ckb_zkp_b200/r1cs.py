"""Host-side R1CS front end: the `zkp_r1cs` interfaces a circuit is written against
(r1cs/src/constraint_system.rs:10-93, r1cs/src/lib.rs:44-187), in Python.

A circuit implements ``generate_constraints(cs)`` (ConstraintSynthesizer, constraint_system.rs:87-93)
and talks to ``cs`` through ``alloc`` / ``alloc_input`` / ``enforce`` with linear combinations
given as ``[(coeff, Variable), ...]``.  Values are plain ints mod r; conversion to Montgomery
limbs happens once, vectorised on the device, when the assignment is handed to the prover.
"""
import numpy as np


class SynthesisError(Exception):
    """r1cs/src/error.rs:7-24"""


class AssignmentMissing(SynthesisError):
    pass


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class Unsatisfiable(SynthesisError):
    pass


class Variable(tuple):
    """Variable(Index): ('in', i) or ('aux', i)  (r1cs/src/lib.rs:44-71)"""
    __slots__ = ()

    def __new__(cls, kind, index):
        return tuple.__new__(cls, (kind, index))


ONE = Variable("in", 0)   # ConstraintSystem::one()


class ProvingAssignment:
    """groth16/src/prover.rs:16-95: records A/B/C rows and the input / aux assignments."""

    def __init__(self, modulus):
        self.p = modulus
        self.input_assignment = []
        self.aux_assignment = []
        self._rows = {"a": ([0], [], []), "b": ([0], [], []), "c": ([0], [], [])}   # row_ptr, (kind, idx), coeff

    def alloc(self, f):
        """`f` is a value or a callable returning one (the reference passes a closure that may fail
        with AssignmentMissing, prover.rs:31-41)."""
        v = f() if callable(f) else f
        if v is None:
            raise AssignmentMissing()
        self.aux_assignment.append(int(v) % self.p)
        return Variable("aux", len(self.aux_assignment) - 1)

    def alloc_input(self, f):
        v = f() if callable(f) else f
        if v is None:
            raise AssignmentMissing()
        self.input_assignment.append(int(v) % self.p)
        return Variable("in", len(self.input_assignment) - 1)

    def enforce(self, a, b, c):
        """push_constraints (groth16/src/lib.rs:128-139) for each of the three linear combinations"""
        for lc, key in ((a, "a"), (b, "b"), (c, "c")):
            ptr, vars_, coeffs = self._rows[key]
            for coeff, var in lc:
                vars_.append(var)
                coeffs.append(int(coeff) % self.p)
            ptr.append(len(vars_))

    @property
    def num_inputs(self):
        return len(self.input_assignment)

    @property
    def num_aux(self):
        return len(self.aux_assignment)

    @property
    def num_constraints(self):
        return len(self._rows["a"][0]) - 1

    def csr(self, which):
        """(row_ptr uint32, col_idx uint32, coeff ints): Input(i) -> i, Aux(i) -> num_inputs + i
        (groth16/src/r1cs_to_qap.rs:34-37)."""
        ptr, vars_, coeffs = self._rows[which]
        ni = self.num_inputs
        cols = np.fromiter((v[1] if v[0] == "in" else ni + v[1] for v in vars_), dtype=np.uint32, count=len(vars_))
        return np.asarray(ptr, dtype=np.uint32), cols, coeffs


def ints_to_limbs(vals, limbs=4):
    """canonical ints -> uint64[n, limbs] little-endian"""
    out = np.zeros((len(vals), limbs), dtype=np.uint64)
    if len(vals) == 0:
        return out
    raw = b"".join(int(v).to_bytes(8 * limbs, "little") for v in vals)
    return np.frombuffer(raw, dtype=np.uint64).reshape(len(vals), limbs).copy()


def limbs_to_int(row):
    return int.from_bytes(np.ascontiguousarray(row, dtype=np.uint64).tobytes(), "little")
