"""CPU oracle (pure-Python big-int restatement) of the ckb-zkp Groth16 prove path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``ckb_zkp_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Parity pin status: the reference (sec-bit/ckb-zkp @ 8f2141a) holds NO golden
vector for this path and its arithmetic lives in un-vendored arkworks 0.2
crates (no Rust toolchain here), so byte-level parity is UNPINNED.  What is
pinned: (1) the reference's own acceptance test -- ``verify_proof == true`` on
the ``Mini`` circuit (groth16/tests/mini.rs:89) -- restated in ``pairing.py``
and ``groth16.verify_proof``; (2) the in-the-exponent Groth16 identity with
known toxic waste; (3) agreement between two independent restatements (this
package and ``oracle/c``).
"""
