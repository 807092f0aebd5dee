"""Static SASS opcode histogram of functions whose mangled name contains a substring.
usage: python tools/sass_mix.py <object|so> <substring> [...]"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if any(k in name for k in sys.argv[2:]):
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f, re.M)
        c = collections.Counter(ops)
        print(name[:120], len(ops))
        print("  " + " ".join(f"{o}:{n}" for o, n in c.most_common(18)))
