"""Host-side hygiene of the native build: every local header a .cu / .cuh of easylp_b200/csrc includes must be a
prerequisite in the Makefile, otherwise `make` keeps a stale object when only the header changes (this happened once in
round 2: bucket_sort.cuh was edited, assemble.o was not rebuilt, and a GPU run measured the old code)."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "easylp_b200", "csrc")


def test_every_local_header_is_a_make_prerequisite():
    with open(os.path.join(CSRC, "Makefile")) as f:
        mk = f.read()
    hdr = re.search(r"^HDR\s*:=\s*(.*)$", mk, re.M).group(1).split()
    listed = {os.path.basename(h) for h in hdr}
    included = set()
    for name in os.listdir(CSRC):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, name)) as f:
                for inc in re.findall(r'^\s*#include\s+"([^"]+)"', f.read(), re.M):
                    included.add(os.path.basename(inc))
    missing = sorted(h for h in included if h not in listed)
    assert not missing, f"headers included but not in the Makefile's HDR list: {missing}"
    # and the objects do depend on that list
    assert re.search(r"^_build/%\.o:\s*%\.cu\s+\$\(HDR\)", mk, re.M)
