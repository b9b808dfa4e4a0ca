"""Byte-level encoders for the C ABI's data formats (include/bp_b200.h): Montgomery
little-endian field elements, 64-byte affine points with (0,0) as the identity."""
R256 = 1 << 256

CURVE_IDS = {"secq256k1": 0, "zorro": 1, "curve25519": 2}
# (base-field modulus, scalar-field modulus)
MODULI = {
    "secq256k1": (0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141, 2**256 - 2**32 - 977),
    "zorro": (57896044618658097711785492504343953927116110621106131396339151912985063395361, 2**255 - 19),
    "curve25519": (2**255 - 19, 2**252 + 27742317777372353535851937790883648493),
}


def enc_fe(v: int, m: int) -> bytes:
    return (v % m * R256 % m).to_bytes(32, "little")


def dec_fe(b: bytes, m: int) -> int:
    return int.from_bytes(b, "little") * pow(R256, -1, m) % m


def enc_scalars(vals, curve: str) -> bytes:
    r = MODULI[curve][1]
    return b"".join(enc_fe(v, r) for v in vals)


def enc_point(P, curve: str) -> bytes:
    q = MODULI[curve][0]
    if P is None:
        return bytes(64)
    return enc_fe(P[0], q) + enc_fe(P[1], q)


def enc_points(pts, curve: str) -> bytes:
    return b"".join(enc_point(P, curve) for P in pts)


def dec_point(b: bytes, curve: str):
    q = MODULI[curve][0]
    x, y = dec_fe(b[:32], q), dec_fe(b[32:64], q)
    return None if (x == 0 and y == 0) else (x, y)
