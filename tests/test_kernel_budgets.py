"""The occupancy and instruction budgets DESIGN.md argues from, checked on the built library without a GPU
(`cuobjdump` reads them out of libb2t.so): the page kernel must fit 8 blocks of 256 threads per SM (32 registers,
<= 28.5 KB of shared memory each; the WordPiece variant with its longer halo 7), the scan kernel 4 blocks (64 registers), and the scan's static instruction count
must not creep up -- it is bound by instruction issue (profiles/k1_experiments_r01.md)."""
import os, re, shutil, subprocess
import pytest
from helpers import ROOT

LIB = os.path.join(ROOT, "tokenizers_b200", "libb2t.so")


def _res():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or libb2t.so not available")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    return res


def test_page_kernel_fits_eight_blocks_per_sm():
    res = _res()
    # SHARED as cuobjdump reports it includes the 1 KB per block the system reserves; an SM has 228 KB
    for model, blocks in ((0, 8), (1, 7)):  # BPE: 8 blocks per SM; WordPiece (416-byte halo): 7
        reg, stack, shared = next(v for k, v in res.items() if f"model_tile_kernelILi{model}E" in k)
        assert reg <= 32, "8 blocks x 256 threads need <= 32 registers per thread (64 K registers per SM)"
        assert shared * blocks <= 233472, f"{blocks} blocks per SM need <= {233472 // blocks} bytes of shared memory each"
        assert stack <= 64


def test_scan_kernel_registers_and_instruction_budget():
    res = _res()
    for kind in (0, 1, 2):
        reg, stack, shared = next(v for k, v in res.items() if f"pretok_scan_kernelILi{kind}ELi256E" in k)
        assert reg <= 64 and shared * 4 <= 233472, "4 blocks x 256 threads per SM"
        assert stack == 0 or kind == 1, "no local memory (the tiktoken variant's rare slow path may keep a few words)"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    body = sass.split("Function : _ZN3b2t18pretok_scan_kernelILi0ELi256E", 1)[1].split("Function : ", 1)[0]
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s", body, flags=re.M))
    assert 1000 < n <= 4300, f"{n} static instructions (round 1: 4096)"
