"""Copies the reference's own example INPUT files used verbatim by the GPU tests into tests/golden/inputs/
(/root/reference does not exist on the GPU box).  These are input decks, not source code; nothing is edited.

    python tests/golden/fetch_reference_inputs.py            # needs /root/reference (or $MPM_REFERENCE)
"""
import os
import shutil

REF = os.environ.get("MPM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["NairnMPM/input/XML_Input/TwoDisks.fmcmd"]

if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, "inputs", os.path.basename(f)))
        print("copied", f)
