#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference learner, vendored by copy at build time.

    python oracle/make_ref.py            # needs /root/reference (the build container only)

Copies ``src/{buffer,agent,model,utils}.py`` of the reference -- byte for byte, nothing edited -- into
``oracle/_ref/src/`` and writes the 12-line ``gymnasium`` stub those modules need at import time
(``src/utils.py:5`` imports gymnasium at module top for its two env wrappers; the real package is not
installed and the hot path never touches it).  ``oracle/_ref/`` is git-ignored (the reference's sources are
never committed) but travels to the GPU box with the snapshot, where ``bench.py --impl reference`` times
``HERBuffer.sample`` + ``DDPG.update`` of exactly these files on the host cores, and
``cpu_baseline.kind`` says "reference".  Test infrastructure only: nothing under goal-conditioned-rl-framework_b200/
imports it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GCRL_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("buffer.py", "agent.py", "model.py", "utils.py")

STUB = '''"""Import-time stand-in for gymnasium (written by oracle/make_ref.py): src/utils.py subclasses two wrappers and
src/env.py (not vendored) builds vector envs; the learner hot path uses none of it."""
import types


class Wrapper:
    def __init__(self, env=None):
        self.env = env


ObservationWrapper = Wrapper
vector = types.SimpleNamespace(AsyncVectorEnv=object, AutoresetMode=types.SimpleNamespace(NEXT_STEP=0))
spaces = types.SimpleNamespace(Dict=dict, Box=object)
'''


def main():
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        print(f"[make_ref] {src} not found: nothing vendored (expected on the GPU box, which uses the prebuilt copy)")
        return 0
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "gymnasium"), exist_ok=True)
    digest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(OUT, "src", f))
        digest[f] = hashlib.sha256(open(os.path.join(OUT, "src", f), "rb").read()).hexdigest()
    open(os.path.join(OUT, "src", "__init__.py"), "w").close()
    with open(os.path.join(OUT, "gymnasium", "__init__.py"), "w") as fh:
        fh.write(STUB)
    with open(os.path.join(OUT, "MANIFEST"), "w") as fh:
        for f, h in digest.items():
            fh.write(f"{h}  src/{f}\n")
    print(f"[make_ref] vendored {len(FILES)} reference files into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
