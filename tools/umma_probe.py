"""Drives tools/libumma_probe.so: which shared-memory layouts does tcgen05.mma kind::tf32 accept for K-major and
MN-major operands?  All layout logic lives here (raw shared-memory images are built with numpy)."""
import ctypes
import sys

import numpy as np
import torch

lib = ctypes.CDLL("tools/libumma_probe.so")
lib.umma_probe.restype = ctypes.c_int
lib.umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint32] * 7 + \
    [ctypes.c_int] * 4 + [ctypes.c_void_p]

NONE, B32ATOM, SW128, SW64, SW32 = 0, 1, 2, 4, 6


def tf32(a):
    return (a.astype(np.float32).view(np.int32) & ~0x1FFF).view(np.float32)


def swz(lin, layout):
    if layout == SW128:
        return lin ^ (((lin >> 7) & 7) << 4)
    if layout == SW64:
        return lin ^ (((lin >> 7) & 3) << 4)
    if layout == SW32:
        return lin ^ (((lin >> 7) & 1) << 4)
    if layout == B32ATOM:
        return lin ^ (((lin >> 7) & 3) << 5)
    return lin


def image(mat, addr_fn, nbytes):
    """mat [R][K] float32 -> raw image; addr_fn(r, k) = byte address"""
    img = np.zeros(nbytes // 4, dtype=np.float32)
    R, K = mat.shape
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    a = addr_fn(r, k)
    assert a.max() < nbytes and len(np.unique(a)) == a.size, "address clash"
    img[a // 4] = mat
    return img


def kmajor_swizzled(rows, K, pitch, layout):
    """K-major: chunks of pitch/4 floats, each chunk [rows][pitch] bytes"""
    kp = pitch // 4
    nb = ((K + kp - 1) // kp) * rows * pitch

    def addr(r, k):
        return swz((k // kp) * rows * pitch + r * pitch + (k % kp) * 4, layout)
    return addr, nb


def mnmajor_rows(nmn, K, pitch, layout):
    """MN-major: [K rows][pitch bytes of MN], MN atoms of pitch/4 floats at stride K*pitch"""
    mp = pitch // 4
    nb = ((nmn + mp - 1) // mp) * K * pitch

    def addr(n, k):
        return swz((n // mp) * K * pitch + k * pitch + (n % mp) * 4, layout)
    return addr, nb


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(name, A, Bm, a_addr, a_nb, b_addr, b_nb, M, N, a_mn, b_mn, a_desc, b_desc, ksteps, a_step, b_step):
    """A [M][K], Bm [N][K]; *_desc = (lbo, sbo, layout)"""
    a_img = torch.from_numpy(image(A, a_addr, a_nb)).cuda()
    b_img = torch.from_numpy(image(Bm, b_addr, b_nb)).cuda()
    out = torch.full((128, 256), float("nan"), device="cuda")
    rc = lib.umma_probe(a_img.data_ptr(), a_nb, b_img.data_ptr(), b_nb, idesc(M, N, a_mn, b_mn), *a_desc, *b_desc,
                        ksteps, a_step, b_step, (N + 15) // 16 * 16, out.data_ptr())
    if rc != 0:
        print(f"{name}: CUDA error {rc}")
        return False
    D = out.cpu().numpy()[:, :N]
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    lanes = np.arange(M) if M == 128 else (np.arange(M) % 16) + 32 * (np.arange(M) // 16)
    got = D[lanes]
    err = np.abs(got - ref).max()
    ok = err < 1e-4 * max(1.0, np.abs(ref).max())
    print(f"{name}: max err {err:.3e} (ref max {np.abs(ref).max():.2f}, got max {np.nanmax(np.abs(got)):.2f}) "
          f"{'OK' if ok else 'FAIL'}")
    if not ok and M == 64:
        alt = D[:64]
        print("    rows-as-lanes-0..63 err", np.abs(alt - ref).max())
    return ok


def main():
    rng = np.random.default_rng(0)
    K = 32
    # T1: known-good, both K-major SW128, M = 128, N = 64
    A = tf32(rng.standard_normal((128, K))); Bm = tf32(rng.standard_normal((64, K)))
    aa, an = kmajor_swizzled(128, K, 128, SW128); ba, bn = kmajor_swizzled(64, K, 128, SW128)
    run("T1 K/K SW128 M=128", A, Bm, aa, an, ba, bn, 128, 64, 0, 0, (16, 1024, SW128), (16, 1024, SW128), 4, 32, 32)
    # T2: M = 64
    A64 = tf32(rng.standard_normal((64, K)))
    aa, an = kmajor_swizzled(64, K, 128, SW128)
    run("T2 K/K SW128 M=64", A64, Bm, aa, an, ba, bn, 64, 64, 0, 0, (16, 1024, SW128), (16, 1024, SW128), 4, 32, 32)
    # MN-major B experiments: B [N][K] with K = 32 "documents", N features
    for N in (32, 64):
        Bn = tf32(rng.standard_normal((N, K)))
        for (lname, layout, sbo) in (("SW128-16B", SW128, 1024), ("128B_ATOM_32B sbo1024", B32ATOM, 1024),
                                     ("128B_ATOM_32B sbo512", B32ATOM, 512)):
            ba2, bn2 = mnmajor_rows(N, K, 128, layout)
            run(f"T3 A K-major SW128 M=64, B MN-major {lname} N={N}", A64, Bn, aa, an, ba2, bn2, 64, N, 0, 1,
                (16, 1024, SW128), (K * 128, sbo, layout), 4, 32, 1024)
        # SWIZZLE_NONE MN-major: [n/4][k][4 floats]

        def none_addr(n, k, K=K):
            return (n // 4) * (K * 16) + k * 16 + (n % 4) * 4
        for (lbo, sbo, tag) in ((128, K * 16, "lbo=kgroup sbo=mngroup"), (K * 16, 128, "lbo=mngroup sbo=kgroup")):
            run(f"T5 B MN-major NONE {tag} N={N}", A64, Bn, aa, an, none_addr, (N // 4) * K * 16, 64, N, 0, 1,
                (16, 1024, SW128), (lbo, sbo, NONE), 4, 32, 128)
    # T8: tail-style SW32 / SW64 MN-major
    for (pitch, layout, N) in ((32, SW32, 8), (64, SW64, 16)):
        Bn = tf32(rng.standard_normal((N, K)))
        ba8, bn8 = mnmajor_rows(N, K, pitch, layout)
        run(f"T8 B MN-major pitch {pitch} N={N}", A64, Bn, aa, an, ba8, bn8, 64, N, 0, 1, (16, 1024, SW128),
            (16, 8 * pitch, layout), 4, 32, 8 * pitch)

    # T6: K-major A with the 32-byte-atom swizzle (what TMA's SWIZZLE_128B_ATOM_32B would leave), M = 128
    aa6, an6 = kmajor_swizzled(128, K, 128, B32ATOM)
    run("T6 A K-major 128B_ATOM_32B M=128", A, Bm, aa6, an6, ba, bn, 128, 64, 0, 0, (16, 1024, B32ATOM),
        (16, 1024, SW128), 4, 32, 32)
    # T7: B K-major 32B atom too
    ba7, bn7 = kmajor_swizzled(64, K, 128, B32ATOM)
    run("T7 B K-major 128B_ATOM_32B", A, Bm, aa, an * 2, ba7, bn7, 128, 64, 0, 0, (16, 1024, SW128),
        (16, 1024, B32ATOM), 4, 32, 32) if False else None


if __name__ == "__main__":
    main()
