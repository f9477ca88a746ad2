"""R1CS front-end: restatement of r1cs/src (ConstraintSystem / LinearCombination)
as far as the prove path needs it, plus the two circuits used everywhere:

* `Mini`  -- groth16/tests/mini.rs:12-44  (x * (y + 2) = z, `num` identical rows)
* `MiMC`  -- the synthetic benchmark circuit of SURVEY.md section 8d, shaped after
  marlin/examples/mimc.rs:26-118 / gadgets/src/hashes/mimc.rs:119-157.

A constraint system here records rows as lists of (coeff, var) with var a global
index: Input(i) -> i, Aux(i) -> num_inputs + i (the mapping applied by
groth16/src/r1cs_to_qap.rs:34-37).  Variables are tagged ('in', i) / ('aux', i)
while synthesising, exactly like r1cs `Index::{Input,Aux}`.
"""
from .fields import stream_field

ONE = ("in", 0)


class ConstraintSystem:
    """ProvingAssignment / KeypairAssembly in one (groth16/src/prover.rs:16-95,
    groth16/src/generator.rs:38-133).  Input 0 is the constant ONE
    (prover.rs:143, generator.rs:158)."""

    def __init__(self, p):
        self.p = p
        self.input_assignment = [1]
        self.aux_assignment = []
        self.at, self.bt, self.ct = [], [], []

    def alloc(self, value):
        self.aux_assignment.append(value % self.p)
        return ("aux", len(self.aux_assignment) - 1)

    def alloc_input(self, value):
        self.input_assignment.append(value % self.p)
        return ("in", len(self.input_assignment) - 1)

    def enforce(self, a, b, c):
        """a, b, c: lists of (coeff, var) -- the LinearCombination terms."""
        for lc, dst in ((a, self.at), (b, self.bt), (c, self.ct)):
            dst.append([(co % self.p, v) for co, v in lc])

    @property
    def num_inputs(self):
        return len(self.input_assignment)

    @property
    def num_aux(self):
        return len(self.aux_assignment)

    @property
    def num_constraints(self):
        return len(self.at)

    def full_assignment(self):
        return self.input_assignment + self.aux_assignment

    def col(self, v):
        return v[1] if v[0] == "in" else self.num_inputs + v[1]

    def rows(self, which):
        src = {"a": self.at, "b": self.bt, "c": self.ct}[which]
        return [[(co, self.col(v)) for co, v in row] for row in src]

    def csr(self, which):
        """(row_ptr, col_idx, coeff) -- the layout crossing the C ABI."""
        ptr, cols, vals = [0], [], []
        for row in self.rows(which):
            for co, cidx in row:
                cols.append(cidx)
                vals.append(co)
            ptr.append(len(cols))
        return ptr, cols, vals

    def is_satisfied(self):
        z, p = self.full_assignment(), self.p
        ev = lambda row: sum(co * z[c] for co, c in row) % p
        for ra, rb, rc in zip(self.rows("a"), self.rows("b"), self.rows("c")):
            if ev(ra) * ev(rb) % p != ev(rc):
                return False
        return True


def mini_circuit(cs, x=2, y=3, z=10, num=10):
    """groth16/tests/mini.rs:18-43."""
    vx = cs.alloc(x)
    vy = cs.alloc(y)
    vz = cs.alloc_input(z)
    for _ in range(num):
        cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])
    return cs


MIMC_SEED = 0x5ECB17


def mimc_circuit(cs, n_constraints, seed=MIMC_SEED):
    """MiMC chain with n_constraints (even) constraints -- SURVEY.md section 8d.
    Stream layout: element 0 = xl0, 1 = xr0, 2 + i = round constant c_i."""
    assert n_constraints % 2 == 0 and n_constraints >= 2
    p = cs.p
    rounds = n_constraints // 2
    xl_v, xr_v = stream_field(seed, 0, p), stream_field(seed, 1, p)
    xl, xr = cs.alloc(xl_v), cs.alloc(xr_v)
    for i in range(rounds):
        c = stream_field(seed, 2 + i, p)
        tmp_v = (xl_v + c) * (xl_v + c) % p
        tmp = cs.alloc(tmp_v)
        cs.enforce([(1, xl), (c, ONE)], [(1, xl), (c, ONE)], [(1, tmp)])
        new_v = ((xl_v + c) * tmp_v + xr_v) % p
        new = cs.alloc_input(new_v) if i == rounds - 1 else cs.alloc(new_v)
        cs.enforce([(1, tmp)], [(1, xl), (c, ONE)], [(1, new), (p - 1, xr)])
        xr, xr_v = xl, xl_v
        xl, xl_v = new, new_v
    return cs
