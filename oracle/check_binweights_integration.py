"""Does the replacement shown in INTEGRATION.md §7 leave the reference's bins unchanged?  TEST INFRASTRUCTURE ONLY.

Build-container check (needs /root/reference; nothing from it is stored in the repo): the source of the unmodified
``Script.utility.match_contigs`` is read at run time, its edge-weight loop nest (the statements from
``for contig in to_bin:`` to the ``edges.append`` inside it, utility.py:466-480) is swapped for the INTEGRATION.md
snippet - with the CPU oracle standing in for the GPU operator, which returns the same bits
(tests/test_binweights.py) - and both versions are run on the same inputs under the same PYTHONHASHSEED.

    PYTHONPATH=/root/reference PYTHONHASHSEED=0 python oracle/check_binweights_integration.py
"""
import inspect
import os
import re
import sys
import textwrap
import types

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
for name in ("igraph", "leidenalg", "Bio", "Bio.SeqIO"):
    sys.modules.setdefault(name, types.ModuleType(name))

from Script import utility as ref_utility                      # noqa: E402
from imputecc_b200 import synth                                 # noqa: E402
from imputecc_b200.binning import bins_to_arrays                # noqa: E402
from oracle import binweights_oracle as bw                      # noqa: E402
from oracle import crwr_oracle as oracle                        # noqa: E402
from oracle.match_inputs import planted_markers, smg_iterations   # noqa: E402

SNIPPET = '''
bin_ptr, members = bins_to_arrays(bins, dict_contigRev)
rows = [dict_contigRev[c] for c in to_bin]
W = bin_contact_weights(normcc_matrix, rows, bin_ptr, members)
for contig in to_bin:
    if contig not in top_nodes:
        top_nodes.append(contig)
edges = [(bins[b][0], contig, W[r, b]) for r, contig in enumerate(to_bin) for b in range(n_bins)]
'''


def patched_match_contigs():
    src = inspect.getsource(ref_utility.match_contigs)
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == "for contig in to_bin:")
    end = next(i for i, l in enumerate(lines) if i > start and "edges.append(" in l)
    indent = re.match(r"\s*", lines[start]).group(0)
    new = lines[:start] + textwrap.indent(SNIPPET.strip("\n"), indent).splitlines() + lines[end + 1:]
    ns = dict(ref_utility.__dict__)
    ns["bins_to_arrays"] = bins_to_arrays
    ns["bin_contact_weights"] = lambda E, rows, ptr, mem: bw.bin_contact_weights(sp.csr_matrix(E), rows, ptr, mem)
    exec(compile("\n".join(new), "<match_contigs with the INTEGRATION.md snippet>", "exec"), ns)
    return ns["match_contigs"], end - start + 1


def main():
    patched, replaced = patched_match_contigs()
    print("replaced %d source lines of the loop nest" % replaced)
    for wl in (synth.Workload("bw1500", 1500, 240, 100, 16), synth.Workload("bw3000dense", 3000, 300, 300, 30)):
        M, idx = synth.make_workload(wl)
        E = sp.csr_matrix(oracle.crwr_impute(M, idx, 0.5, 80, dtype=np.float32))
        names = ["contig%05d" % i for i in idx]
        rev = {nm: k for k, nm in enumerate(names)}
        contig_markers = planted_markers(wl, idx, np.random.default_rng(2))
        smg = smg_iterations(contig_markers)
        data = E.tocoo().data
        args = (smg, contig_markers, E.tolil(), rev, np.percentile(data, 50), np.percentile(data, 0))
        want = ref_utility.match_contigs(*args)
        got = patched(*args)
        same = want[0] == got[0] and want[1] == got[1] and want[2] == got[2]
        print("%s: %d bins, identical bins / bin_of_contig / n_bins: %s" % (wl.name, len(want[0]), same))
        assert same


if __name__ == "__main__":
    main()
