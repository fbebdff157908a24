"""Prints the hottest SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv` output."""
import csv
import sys


def sections(path):
  rows = list(csv.reader(open(path)))
  cur = None
  for r in rows:
    if r and r[0] == "Kernel Name":
      cur = {"name": r[1], "hdr": None, "rows": []}
      yield cur
    elif cur is not None and cur["hdr"] is None:
      cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
      cur["rows"].append(r)


def main(path, frac=0.004, only=None):
  for sec in list(sections(path)):
    if only and only not in sec["name"]:
      continue
    h = sec["hdr"]
    isrc, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in sec["rows"])
    smp = sum(int(r[iss]) for r in sec["rows"])
    print("=====", sec["name"][:100], "total warp-inst %.1fM samples %d" % (tot / 1e6, smp))
    for k, r in enumerate(sec["rows"]):
      c = int(r[ie])
      if c > tot * frac or int(r[iss]) > smp * frac * 2:
        print(f"{k:4d} {c/1e6:8.1f}M {100*c/tot:5.1f}% smp={100*int(r[iss])/max(smp,1):5.1f}% {r[isrc].strip()[:100]}")


if __name__ == "__main__":
  main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.004, sys.argv[3] if len(sys.argv) > 3 else None)
