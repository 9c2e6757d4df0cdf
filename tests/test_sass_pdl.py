"""Build-time guard for the programmatic-dependent-launch chain (csrc/launch.cuh): no kernel may read global memory that an
earlier kernel of the chain writes BEFORE its griddepcontrol.wait (ACQBULK in SASS).  The compiler is free to hoist __ldg /
const __restrict__ loads above the wait's asm statement; it did so once with the attention kernel's item count, which then
used the count of the previous pass.  Allowed ahead of the wait: loads of load-time constants (biases, resize taps).
Runs on CPU: cuobjdump disassembles the sm_100a library that build() produced."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import PKG

LIB = os.path.join(PKG, "lib", "libvcg_b200.so")
# kernels whose pre-wait global loads are constants written once at load / set_frame_size time
CONSTANT_PROLOGUE = ("conv23h_kernel", "conv23h2_kernel", "resize_preprocess_u8_kernel")


@pytest.fixture(scope="module")
def sass():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    return subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True).stdout


def test_no_global_load_ahead_of_the_pdl_wait(sass):
    fn, waited, offenders, with_wait = None, False, {}, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, waited = m.group(1), False
            continue
        if fn is None or "/*" not in line:
            continue
        if "ACQBULK" in line:
            if not waited:
                with_wait += 1
            waited = True
        elif not waited and re.search(r"\b(LDG|STG|ATOMG|REDG|RED\.E|UTMALDG|UTMASTG)\b", line):
            offenders.setdefault(fn, []).append(line.strip()[:80])
    assert with_wait >= 40, f"only {with_wait} kernels with a griddepcontrol.wait found: wrong library?"
    # kernels that never wait are not part of the chain (weight packing at load time, the debug checksum): fine
    real = {}
    for f, lines in offenders.items():
        body = sass.split("Function : " + f, 1)[1].split("Function : ", 1)[0]
        if "ACQBULK" in body and not any(name in f for name in CONSTANT_PROLOGUE):
            real[f] = lines[:3]
    assert not real, f"global memory access ahead of griddepcontrol.wait: {real}"


def test_every_kernel_launched_with_the_attribute_waits(sass):
    """A kernel launched with programmatic stream serialization that never executes griddepcontrol.wait could finish
    before its predecessor and break the chain for everything behind it: every kernel handed to launch_pdl must wait."""
    csrc = os.path.join(PKG, "csrc")
    names = set()
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".h")):
            text = open(os.path.join(csrc, f)).read()
            names.update(re.findall(r"launch_pdl\(\s*([A-Za-z_0-9]+_kernel)\b", text))
    assert len(names) >= 30, names
    bodies = sass.split("Function : ")[1:]
    checked = 0
    for body in bodies:
        fn = body.split("\n", 1)[0].strip()
        if any(re.search(r"\d+" + n + r"(I|E|P|v)", fn) for n in names):   # Itanium mangling: <len><name>
            checked += 1
            assert "ACQBULK" in body, f"{fn} is launched with the dependent-launch attribute but never waits"
    assert checked >= len(names), (checked, len(names))


def test_library_carries_tcgen05_and_tma_code(sass):
    """The hot kernels are tcgen05 / TMEM / TMA code for sm_100a, not recompiled mma.sync: the mnemonics must be there."""
    assert "arch = sm_100a" in sass
    for mnemonic, at_least in (("UTCHMMA", 50), ("LDTM", 30), ("STTM", 2), ("UTMALDG", 50), ("UTMASTG", 5), ("UTCBAR", 10)):
        n = len(re.findall(r"\b" + mnemonic + r"\b", sass))
        assert n >= at_least, f"{mnemonic}: {n} occurrences"
