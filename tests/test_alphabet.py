"""The code tables of convert_msa_to_internal, after the reference's alphabet_utest (lib/CMakeLists.txt:249-258):
kb200_alphabet (the tables the device encoder is built from) against the reference's create_alphabet
(lib/src/alphabet.c:140) for the three alphabets of the path -- live where oracle/_ref exists, and against the tables
committed in tests/golden/alphabet.npz (written by this file's gen() from the reference)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # `python tests/test_alphabet.py` regenerates
import kbind  # noqa: E402
from kalign_b200 import _lib  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alphabet.npz")
ALPHABETS = (5, 13, 23)          # ALPHA_defDNA, ALPHA_redPROTEIN, ALPHA_ambigiousPROTEIN (alphabet.h:18-22)


class RefAlphabet(C.Structure):
    _fields_ = [("to_internal", C.c_int8 * 128), ("to_external", C.c_int8 * 32), ("type", C.c_int), ("L", C.c_int)]


def ref_table(letters):
    ref = kbind.ref()
    ref.create_alphabet.restype = C.POINTER(RefAlphabet)
    ref.create_alphabet.argtypes = [C.c_int]
    a = ref.create_alphabet(letters)
    t = np.array(a.contents.to_internal[:], dtype=np.int8)
    L = int(a.contents.L)
    C.CDLL(None).free(a)
    return t, L


def gen():
    rec = {}
    for n in ALPHABETS:
        t, L = ref_table(n)
        rec["t%d" % n] = t
        rec["L%d" % n] = L
    np.savez_compressed(G, **rec)


@pytest.mark.parametrize("letters", ALPHABETS)
def test_tables_reproduce_golden(letters):
    z = np.load(G)
    t, L = _lib.alphabet(letters)
    assert np.array_equal(t, z["t%d" % letters]) and L == int(z["L%d" % letters])
    # lower case maps like upper case, everything that is not a letter is outside the alphabet
    for ch in range(128):
        if chr(ch).isalpha():
            assert t[ch] == t[ord(chr(ch).upper())]
        else:
            assert t[ch] == -1
    assert t.max() == L - 1
    with pytest.raises(RuntimeError):
        _lib.alphabet(21)


@pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("letters", ALPHABETS)
def test_tables_equal_reference_live(letters):
    t, L = _lib.alphabet(letters)
    rt, rL = ref_table(letters)
    assert np.array_equal(t, rt) and L == rL


if __name__ == "__main__":
    gen()
