"""Radix-2 NTT: restatement of ark-poly 0.2 `Radix2EvaluationDomain`.

Not in /root/reference (un-vendored `ark-poly = "0.2"`); anchored on the call
sites groth16/src/r1cs_to_qap.rs:123-169.  Published behaviour: size = next
power of two, group_gen = get_root_of_unity(size); fft = bit-reversal +
Cooley-Tukey DIT, natural order in and out; ifft = fft with group_gen^-1 then
scale by size^-1; coset_fft = distribute_powers(g) then fft; coset_ifft = ifft
then distribute_powers(g^-1); g = Fr::multiplicative_generator().
"""


class Domain:
    def __init__(self, fr, min_size):
        """EvaluationDomain::new(min_size); raises ValueError when too large
        (reference maps None -> SynthesisError::PolynomialDegreeTooLarge)."""
        size = 1
        log = 0
        while size < min_size:
            size <<= 1
            log += 1
        if log > fr.two_adicity:
            raise ValueError("PolynomialDegreeTooLarge")
        self.fr, self.p = fr, fr.p
        self.size, self.log = size, log
        self.group_gen = fr.root_of_unity(log)
        self.group_gen_inv = pow(self.group_gen, -1, self.p)
        self.size_inv = pow(size, -1, self.p)
        self.g = fr.generator
        self.g_inv = pow(self.g, -1, self.p)

    # -- core ---------------------------------------------------------------
    def _fft(self, a, omega):
        n, p, log = self.size, self.p, self.log
        a = list(a) + [0] * (n - len(a))
        for k in range(n):
            rk = int(format(k, "0%db" % log)[::-1], 2) if log else 0
            if k < rk:
                a[k], a[rk] = a[rk], a[k]
        m = 1
        for _ in range(log):
            w_m = pow(omega, n // (2 * m), p)
            for k in range(0, n, 2 * m):
                w = 1
                for j in range(m):
                    t = a[k + j + m] * w % p
                    a[k + j + m] = (a[k + j] - t) % p
                    a[k + j] = (a[k + j] + t) % p
                    w = w * w_m % p
            m *= 2
        return a

    def fft(self, a):
        return self._fft(a, self.group_gen)

    def ifft(self, a):
        return [x * self.size_inv % self.p for x in self._fft(a, self.group_gen_inv)]

    def _distribute(self, a, g):
        out, pw = [], 1
        for x in a:
            out.append(x * pw % self.p)
            pw = pw * g % self.p
        return out

    def coset_fft(self, a):
        a = list(a) + [0] * (self.size - len(a))
        return self.fft(self._distribute(a, self.g))

    def coset_ifft(self, a):
        return self._distribute(self.ifft(a), self.g_inv)

    def vanishing_at(self, x):
        return (pow(x, self.size, self.p) - 1) % self.p

    def lagrange_coeffs_at(self, tau):
        """evaluate_all_lagrange_coefficients(tau) (tau outside the domain)."""
        p, n = self.p, self.size
        z = self.vanishing_at(tau)
        assert z != 0
        out, w = [], 1
        for _ in range(n):
            # L_i(tau) = Z(tau) * w^i / (n * (tau - w^i))
            out.append(z * w % p * pow(n * (tau - w) % p, -1, p) % p)
            w = w * self.group_gen % p
        return out


def dft_naive(a, omega, p):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, p) for j in range(n)) % p for i in range(n)]
