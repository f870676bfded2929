"""BN254 scalar field (acir_field::FieldElement) on Python ints.  TEST INFRASTRUCTURE ONLY.

Follows acir_field/src/generic_ark.rs (reference file:line cited per function).  Values are
always canonical ints in [0, P).  The arithmetic itself is ark-ff 0.4.2 / ark-bn254 0.4.0
(crates.io, not vendored in the reference): pure modular arithmetic, restated with big ints.
"""

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
MAX_NUM_BITS = 254  # generic_ark.rs:181-183, test :440-443


def add(a, b):  # generic_ark.rs:383-394
    return (a + b) % P


def sub(a, b):  # generic_ark.rs:396-406
    return (a - b) % P


def mul(a, b):  # generic_ark.rs:367-373
    return (a * b) % P


def neg(a):  # generic_ark.rs:360-365
    return (-a) % P


def inverse(a):  # generic_ark.rs:242-245 -- inverse of zero is zero
    return pow(a, P - 2, P) if a % P else 0


def div(a, b):  # generic_ark.rs:375-380 -- self * rhs.inverse()
    return mul(a, inverse(b))


def to_be_bytes(a):  # generic_ark.rs:269-277
    return int(a).to_bytes(32, "big")


def to_hex(a):  # generic_ark.rs:257-262
    return to_be_bytes(a).hex()


def from_be_bytes_reduce(b):  # generic_ark.rs:281-283
    return int.from_bytes(bytes(b), "big") % P


def from_hex(s):  # generic_ark.rs:263-267
    s = s[2:] if s.startswith("0x") else s
    return from_be_bytes_reduce(bytes.fromhex(s))


def num_bits(a):  # generic_ark.rs:214-221 -- bit length of the canonical value, 0 for 0
    return int(a).bit_length()


def fetch_nearest_bytes(a, nbits):  # generic_ark.rs:305-317 -- low ceil(nbits/8) bytes, little-endian
    n = ((nbits + 7) // 8)
    le = to_be_bytes(a)[::-1]
    if n > 32:
        raise IndexError("fetch_nearest_bytes: more than 32 bytes requested (reference panics)")
    return le[:n]


def _mask(a, nbits):  # generic_ark.rs:322-326 + mask_vector_le :446-473
    if nbits >= 256:
        return a
    return a & ((1 << nbits) - 1)


def and_(a, b, nbits):  # generic_ark.rs:328-354
    return (_mask(a, nbits) & _mask(b, nbits)) % P


def xor(a, b, nbits):  # generic_ark.rs:328-354
    return (_mask(a, nbits) ^ _mask(b, nbits)) % P


def to_u128(a):  # generic_ark.rs:227-230 -- low 128 bits
    return a & ((1 << 128) - 1)
