"""Parses the reference's public enum header ($PASTIX_REFERENCE/src/common/src/api.h:
IPARM_*, DPARM_*, API_* values) into pastix_b200/lib/api_enums.json so that the Python binding
of the drop-in library (pastix_b200/pastix_api.py) addresses iparm/dparm by the reference's own
names.  Generated next to the library at build time, never committed."""
import json, os, re, sys

def main():
    ref = os.environ.get("PASTIX_REFERENCE", "/root/reference")
    src = os.path.join(ref, "src/common/src/api.h")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib", "api_enums.json")
    if not os.path.exists(src):
        print("reference api.h not found; keeping", out)
        return 0
    txt = open(src, encoding="latin-1").read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    enums = {}
    for m in re.finditer(r"^\s*((?:IPARM|DPARM|API|MODULE|ERR)_[A-Z0-9_a-z]+)\s*=\s*(-?\d+)", txt, flags=re.M):
        enums[m.group(1)] = int(m.group(2))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(enums, open(out, "w"), indent=0, sort_keys=True)
    print("wrote", out, len(enums), "names")
    return 0

if __name__ == "__main__":
    sys.exit(main())
