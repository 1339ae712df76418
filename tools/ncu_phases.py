"""tools/ncu_phases.py -- per-phase instruction / lane / stall-sample shares of the envelope kernels from the source
counters of an `ncu --set full --import-source on` report: the SASS stream is cut at the CTA / warp barriers
(staging | TMA wait | build | seam round 0 | seam rounds 1+ | query), which is where the phases of band_hull.cuh end.
usage: python tools/ncu_phases.py gpurun_out/r2_prof.ncu-rep >> profiles/r2_ncu_summary.md"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
print("\n## phases of the envelope kernels (source counters of the same capture, tools/ncu_phases.py)\n")
print("Segments of the SASS stream between barriers; `lanes` = thread instructions / warp instructions; `samples` = share of all\n"
      "warp-state samples of the kernel taken inside the segment (a warp parked at a barrier is sampled at the barrier's segment end).\n")
seen = set()
for k in kernels:
    if "band_strided" not in k["name"] or k["name"] in seen:
        continue
    seen.add(k["name"])
    h = k["hdr"]; ia, ii, it, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ii]) for r in k["rows"]); tots = sum(int(r[iss]) for r in k["rows"]) or 1
    print(f"### {k['name'].split('(int *')[0]} — {tot / 1e9:.3f} G warp instructions\n\n| segment (ends with) | warp instructions | share | lanes | samples |\n|---|---|---|---|---|")
    bars = [i for i, r in enumerate(k["rows"]) if "BAR.SYNC" in r[ia] or "WARPSYNC" in r[ia]]
    prev = 0
    for bi in bars + [len(k["rows"]) - 1]:
        seg = k["rows"][prev:bi + 1]
        si = sum(int(r[ii]) for r in seg); st = sum(int(r[it]) for r in seg); ss = sum(int(r[iss]) for r in seg)
        if si:
            print(f"| SASS [{prev}, {bi}] {k['rows'][bi][ia].strip().split()[0]} | {si / 1e6:.0f} M | {100 * si / tot:.1f} % | {st / si:.1f} | {100 * ss / tots:.1f} % |")
        prev = bi + 1
    print()

# line-per-thread kernel: no barriers to cut at -- totals and the SASS lines where warps wait the most
seen = set()
for k in kernels:
    if "line_hull" not in k["name"] or k["name"] in seen:
        continue
    seen.add(k["name"])
    h = k["hdr"]; ia, ii, it, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ii]) for r in k["rows"]); tott = sum(int(r[it]) for r in k["rows"]); tots = sum(int(r[iss]) for r in k["rows"]) or 1
    print(f"### {k['name'].split('(const int *')[0]} — {tot / 1e9:.3f} G warp instructions, {tott / max(tot, 1):.1f} lanes\n")
    print("| SASS line | instruction | executed | lanes | samples |\n|---|---|---|---|---|")
    top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iss]))[:8]
    for i in sorted(top):
        r = k["rows"][i]
        ex = int(r[ii])
        print(f"| {i} | `{r[ia].strip()[:60]}` | {ex / 1e6:.1f} M | {int(r[it]) / max(ex, 1):.0f} | {100 * int(r[iss]) / tots:.1f} % |")
    print()
