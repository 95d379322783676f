#!/usr/bin/env python
"""CPU model of the register-level data movement of the symmetric packed ICE kernel (hc_ice_sym.cu):
ldmatrix (plain and .trans, b16 elements), the two PRMT selectors, and mma.sync.m16n8k32.u8.u8.s32 fragment
layouts, applied to the tile layout the pack kernel writes.  It checks that one pass over a 16x32 tile pair
yields BOTH the row sums (tile x bias planes of the columns) and the column sums (tile^T x bias planes of the
rows).  Run:  python tools/sym_fragment_model.py   (prints OK)."""
import numpy as np

rng = np.random.default_rng(0)


def tile_bytes(T):
    """16x32 u8 tile -> 512 bytes: four 8x8 b16 matrices (h, k) in the order (0,0) (1,0) (0,1) (1,1); matrix row i is
    the 16 bytes of tile row 8h+i, byte columns 16k .. 16k+15."""
    out = np.zeros(512, np.uint8)
    for m, (h, k) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
        for i in range(8):
            out[128 * m + 16 * i:128 * m + 16 * i + 16] = T[8 * h + i, 16 * k:16 * k + 16]
    return out


def ldmatrix_x4(smem, addr_of_lane, trans):
    """ldmatrix.sync.aligned.m8n8.x4[.trans].shared.b16: lane l supplies the address of row l%8 of matrix l/8.
    Returns regs[lane][j] as 4 bytes (little endian)."""
    regs = np.zeros((32, 4, 4), np.uint8)
    for j in range(4):
        M = np.zeros((8, 8, 2), np.uint8)           # 8 rows x 8 b16 elements
        for i in range(8):
            a = addr_of_lane[8 * j + i]
            M[i] = smem[a:a + 16].reshape(8, 2)
        for l in range(32):
            g, q = l // 4, l % 4
            if not trans:
                regs[l, j, 0:2] = M[g, 2 * q]; regs[l, j, 2:4] = M[g, 2 * q + 1]
            else:                                   # element [row g][col 2q, 2q+1] of M^T = M[2q][g], M[2q+1][g]
                regs[l, j, 0:2] = M[2 * q, g]; regs[l, j, 2:4] = M[2 * q + 1, g]
    return regs


def prmt(a, b, sel):
    """prmt.b32 d, a, b, sel (default mode): byte i of d = byte sel_nibble_i of {b:a}."""
    src = np.concatenate([a, b])
    return np.array([src[(sel >> (4 * i)) & 7] for i in range(4)], np.uint8)


def mma_m16n8k32(A_frag, B_frag):
    """A_frag[lane][4 regs][4 bytes], B_frag[lane][2 regs][4 bytes] -> C[lane][4] int32 with the PTX layouts:
    A (row): a0 (g, 4q..), a1 (g+8, 4q..), a2 (g, 16+4q..), a3 (g+8, 16+4q..); B (col): b0 k=4q.. n=g, b1 k=16+4q..;
    C: c0 (g, 2q) c1 (g, 2q+1) c2 (g+8, 2q) c3 (g+8, 2q+1)."""
    A = np.zeros((16, 32), np.int64); B = np.zeros((32, 8), np.int64)
    for l in range(32):
        g, q = l // 4, l % 4
        A[g, 4 * q:4 * q + 4] = A_frag[l, 0]; A[g + 8, 4 * q:4 * q + 4] = A_frag[l, 1]
        A[g, 16 + 4 * q:20 + 4 * q] = A_frag[l, 2]; A[g + 8, 16 + 4 * q:20 + 4 * q] = A_frag[l, 3]
        B[4 * q:4 * q + 4, g] = B_frag[l, 0]; B[16 + 4 * q:20 + 4 * q, g] = B_frag[l, 1]
    D = A @ B
    C = np.zeros((32, 4), np.int64)
    for l in range(32):
        g, q = l // 4, l % 4
        C[l] = [D[g, 2 * q], D[g, 2 * q + 1], D[g + 8, 2 * q], D[g + 8, 2 * q + 1]]
    return C


ROWPERM = lambda q, j: (2 * q, 2 * q + 1, 8 + 2 * q, 9 + 2 * q)[j]      # k-slot 4q+j of product 2 <-> tile row


def dig1_frag(dcols):
    """B fragment of product 1: dcols[32 columns][8 planes] -> [lane][2][4]"""
    f = np.zeros((32, 2, 4), np.uint8)
    for l in range(32):
        g, q = l // 4, l % 4
        f[l, 0] = dcols[4 * q:4 * q + 4, g]; f[l, 1] = dcols[16 + 4 * q:20 + 4 * q, g]
    return f


def dig2_frag(drows_a, drows_b):
    """B fragment of product 2 for a strip pair: rows of strip a in k-slots 0..15, strip b in 16..31, permuted"""
    f = np.zeros((32, 2, 4), np.uint8)
    for l in range(32):
        g, q = l // 4, l % 4
        f[l, 0] = [drows_a[ROWPERM(q, j), g] for j in range(4)]
        f[l, 1] = [drows_b[ROWPERM(q, j), g] for j in range(4)]
    return f


def main():
    Ta = rng.integers(0, 256, (16, 32)).astype(np.uint8)      # strip a
    Tb = rng.integers(0, 256, (16, 32)).astype(np.uint8)      # strip b (16 rows below)
    dcols = rng.integers(0, 256, (32, 8)).astype(np.uint8)    # byte planes of the bias of the 32 columns
    drows = rng.integers(0, 256, (32, 8)).astype(np.uint8)    # ... of the 32 rows (a then b)
    smem = np.concatenate([tile_bytes(Ta), tile_bytes(Tb)])
    # ---- product 1: rows x planes, per tile: ldmatrix.x4 with lane address = tile + 16 * lane
    for base, T in ((0, Ta), (512, Tb)):
        regs = ldmatrix_x4(smem, [base + 16 * l for l in range(32)], trans=False)
        C = mma_m16n8k32(regs, dig1_frag(dcols))
        ref = T.astype(np.int64) @ dcols.astype(np.int64)      # [16 rows][8 planes]
        for l in range(32):
            g, q = l // 4, l % 4
            assert list(C[l]) == [ref[g, 2 * q], ref[g, 2 * q + 1], ref[g + 8, 2 * q], ref[g + 8, 2 * q + 1]]
    # ---- product 2: columns x planes over the 32 rows of the pair, one MMA per 16-column half k
    ref2 = np.concatenate([Ta, Tb]).astype(np.int64).T @ drows.astype(np.int64)     # [32 cols][8 planes]
    for k in range(2):
        # matrices: A(0,k) A(1,k) B(0,k) B(1,k); in tile_bytes order (0,0)=0 (1,0)=1 (0,1)=2 (1,1)=3
        mats = [0 + 2 * k, 1 + 2 * k]
        addr = [128 * mats[0] + 16 * i for i in range(8)] + [128 * mats[1] + 16 * i for i in range(8)] + \
               [512 + 128 * mats[0] + 16 * i for i in range(8)] + [512 + 128 * mats[1] + 16 * i for i in range(8)]
        r = ldmatrix_x4(smem, addr, trans=True)
        frag = np.zeros((32, 4, 4), np.uint8)
        for l in range(32):
            frag[l, 0] = prmt(r[l, 0], r[l, 1], 0x6420)     # X_a: column 2g,   rows {2q, 2q+1, 8+2q, 9+2q} of strip a
            frag[l, 1] = prmt(r[l, 0], r[l, 1], 0x7531)     # Y_a: column 2g+1
            frag[l, 2] = prmt(r[l, 2], r[l, 3], 0x6420)     # X_b
            frag[l, 3] = prmt(r[l, 2], r[l, 3], 0x7531)     # Y_b
        C = mma_m16n8k32(frag, dig2_frag(drows[:16], drows[16:]))
        for l in range(32):
            g, q = l // 4, l % 4
            c_lo, c_hi = 16 * k + 2 * g, 16 * k + 2 * g + 1     # m = g <-> column 2g, m = g+8 <-> column 2g+1
            assert list(C[l]) == [ref2[c_lo, 2 * q], ref2[c_lo, 2 * q + 1], ref2[c_hi, 2 * q], ref2[c_hi, 2 * q + 1]], (k, l)
    print("OK: one pass over a tile pair yields row sums and column sums")


if __name__ == "__main__":
    main()
