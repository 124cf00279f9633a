"""TEST INFRASTRUCTURE: builds (once) and loads the CPU-emulated build of the kernel sources (tests/emu)."""
import os
import subprocess

from multimodalgame_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libmmg_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "multimodalgame_b200", "csrc")
_lib = None


def emu_library():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(os.path.join(d, f)) for d in (CSRC, EMU_DIR) for f in os.listdir(d)
                     if not f.endswith(".so"))
        if not os.path.isfile(EMU_LIB) or os.path.getmtime(EMU_LIB) < newest:
            subprocess.check_call(["bash", os.path.join(EMU_DIR, "build_emu.sh")], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
        _lib = capi.Library(EMU_LIB)
    return _lib
