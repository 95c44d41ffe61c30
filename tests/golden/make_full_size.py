"""Generates tests/golden/full_size.json: fingerprints of what the ORACLE (oracle/, the CPU restatement of the reference,
pinned on the reference's own tests) produces at BASELINE.json's full sizes, for the workloads of SURVEY.md 8d:

  C3  1 000 001 changes, 10 actors (headline)      C4_100k  100k set ops, 10 000 keys, 100 actors (Lamport-conflict heavy)
  C2  100 001 changes, one actor                   C2b the 100k inserts of C2 in one change

For each: SHA-256 of save() (canonical: every row, succ list and change record), heads, maxOp, and the digests of the
incremental patch of applyChanges(init(), all changes) and of getPatch() (canonical JSON: keys sorted, edits in order -
tests/parity_checks.py patch_digest). The oracle needs minutes for these sizes (it keeps the reference's 600-op blocks and
is super-linear in document length), so its results are committed; the GPU tests (tests/test_zz_full_size.py) recompute the
same digests from the CUDA engine's output.

    python tests/golden/make_full_size.py [C3 C4 C2 C2b]        (run in the build container; writes next to itself)
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, '..'), os.path.join(HERE, '..', '..')]
import oracle                                   # noqa: E402
import parity_checks                            # noqa: E402
from automerge_classic_b200 import tracegen     # noqa: E402

# C4 at its full size (1M ops, 100 actors) is out of the oracle's reach: like the reference it seeks through the op blocks
# linearly, 50k ops take 43 s, 100k ops 172 s (quadratic: about 5 hours for 1M). Its fingerprint is taken at 100k ops with the
# full 100 actors / 10 000 keys ('C4_100k'); the 1M-op document is covered by size-independent properties on the GPU.
SIZES = {'C3': ('C3', 1000000, 10), 'C4_100k': ('C4', 100000, 100), 'C2': ('C2', 100000, 1), 'C2b': ('C2b', 100000, 1)}


def main():
    out_path = os.path.join(HERE, 'full_size.json')
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for cfg in (sys.argv[1:] or list(SIZES)):
        gen, n, a = SIZES[cfg]
        t = tracegen.generate(gen, n, a)
        o = oracle.OracleDoc()
        t0 = time.time()
        patch = o.apply_changes(t.changes())
        dt = time.time() - t0
        s = o.save()
        res[cfg] = {'config': gen, 'ops_requested': n, 'n_ops': t.n_ops, 'n_changes': t.n_changes, 'n_actors': a, 'change_bytes': int(t.offsets[-1]),
                    'save_bytes': len(s), 'save_sha256': hashlib.sha256(s).hexdigest(), 'heads': o.heads(), 'max_op': o.max_op(),
                    'patch_sha256': parity_checks.patch_digest(patch), 'get_patch_sha256': parity_checks.patch_digest(o.get_patch()),
                    'oracle_seconds': round(dt, 1)}
        print(cfg, res[cfg], flush=True)
        json.dump(res, open(out_path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
