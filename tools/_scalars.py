"""Synthetic MSM scalars for the tools: uniform field elements of BLS12-381 Fr (what ring columns and polynomial coefficients look
like).  Bit 254 is set for ~55 % of them, which decides how the top window of a signed-digit recoding is populated - a plain
254-bit mask hides that (and made c = 15 look as good as c = 16 at 2^17)."""
import numpy as np

R_BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def fr_uniform(rng, m):
    raw = rng.integers(0, 256, size=(m, 40), dtype=np.uint8)
    return np.frombuffer(b"".join((int.from_bytes(r.tobytes(), "little") % R_BLS).to_bytes(32, "little") for r in raw), np.uint8).reshape(m, 32).copy()
