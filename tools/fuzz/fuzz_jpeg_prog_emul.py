"""Fuzzer of the progressive scan decoder compiled for the host (see run.sh): python fuzz_jpeg_prog_emul.py <libjpegprog_asan.so> <seed> <n>.
Mutated progressive streams go through the planner and through prog_decode_scan / prog_dc_difference -- the code the CUDA kernels run --
under AddressSanitizer + UndefinedBehaviorSanitizer.  Every call must return (0 or a status); a finding is a sanitizer report."""
import ctypes as C, numpy as np, sys, cv2
L = C.CDLL(sys.argv[1])
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
def synth(h, w, q=90, sub=None, rst=0, gray=False):
    img = rng.integers(0, 255, (h, w) if gray else (h, w, 3)).astype(np.uint8)
    p = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    if sub is not None: p += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sub]
    if rst: p += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    return bytearray(cv2.imencode(".jpg", img, p)[1].tobytes())
seeds = [synth(33, 47), synth(64, 64, sub=cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444), synth(40, 24, rst=2), synth(17, 19, gray=True),
         synth(48, 80, rst=1, sub=cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422), synth(24, 24, q=30, sub=cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411)]
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
from test_jpeg_prog_cpu import to_multiscan_baseline          # sequential frames coded in several scans (test encoder)
for rst in (0, 3):
    img = rng.integers(0, 255, (40, 56, 3)).astype(np.uint8)
    seeds.append(bytearray(to_multiscan_baseline(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes(), rst)[0]))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
hist = {}
info = (C.c_int * 8)()
for it in range(n):
    s = bytearray(seeds[int(rng.integers(0, len(seeds)))])
    for _ in range(int(rng.integers(1, 3))):
        k = int(rng.integers(0, 7))
        if k == 0:
            for _ in range(rng.integers(1, 8)): s[rng.integers(0, len(s))] = rng.integers(0, 256)
        elif k == 1:
            s = s[:rng.integers(4, len(s))]
        elif k == 2:      # corrupt a scan header (Ss, Se, Ah/Al, table ids) or a DHT
            pos = [i for i in range(2, len(s) - 12) if s[i] == 0xFF and s[i + 1] in (0xDA, 0xC4, 0xC2, 0xDD)]
            if pos:
                p = pos[rng.integers(0, len(pos))]; s[p + 4 + int(rng.integers(0, 8))] = rng.integers(0, 256)
        elif k == 3:      # markers inside the entropy data
            for _ in range(rng.integers(1, 5)):
                p = rng.integers(len(s) // 3, len(s) - 1); s[p] = 0xFF; s[p + 1] = int(rng.choice([0xD0, 0xD3, 0xD7, 0xD9, 0x00, 0xFF, 0xC4, 0xDA]))
        elif k == 4:
            p = rng.integers(0, len(s)); s[p:p] = bytes(rng.integers(0, 256, rng.integers(1, 9)).astype(np.uint8))
        elif k == 5:
            p = rng.integers(0, len(s)); del s[p:p + rng.integers(1, 9)]
        # k == 6: leave it
        if len(s) < 4: s = bytearray(seeds[0])
    buf = (C.c_uint8 * len(s)).from_buffer_copy(bytes(s))
    rc = L.emul_jpeg_progressive(buf, C.c_size_t(len(s)), None, info)
    if rc == 0:
        ncomp, mcux, mcuy, bpm = info[0], info[1], info[2], info[3]
        if 0 < mcux * mcuy * bpm <= 1 << 20:
            coef = np.zeros((mcux * mcuy * bpm, 64), np.int16)
            rc = L.emul_jpeg_progressive(buf, C.c_size_t(len(s)), coef.ctypes.data_as(C.c_void_p), info)
            rc = ("decoded", rc, info[6])
    hist[rc] = hist.get(rc, 0) + 1
print("seed", sys.argv[2] if len(sys.argv) > 2 else 0, "iterations", n, "status histogram", hist, "- no sanitizer report")
