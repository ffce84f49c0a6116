"""Builds (once) and locates the host-emulation test harness tests/emu/_build/libibis_b200_emu.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_LIB = os.path.join(HERE, "emu", "_build", "libibis_b200_emu.so")


def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu")], check=True)
    return EMU_LIB


def emu_host_lib():
    emu_lib()
    return os.path.join(HERE, "emu", "_build", "libibis_b200_host_emu.so")
