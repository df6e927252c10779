#!/usr/bin/env python
"""Classify a compute-sanitizer log of tools/sanitize_small.py: print a summary, exit 1 on any UNLISTED report.

Allow-list (one class, racecheck only): a hazard whose write side has no instruction inside the kernel (the tool prints
the kernel name with an offset of 0xffff...: the tcgen05.alloc.cta_group::2 result landing in shared memory) and whose
read side is the load of the TMEM base address in a CTA-pair kernel (k_conv3x3_tc2 / k_conv3x3_tc4).  The same report
fires on tools/sanitizer_repro/tmem_alloc_pair.cu, which contains only the documented allocation hand-off."""
import re
import sys

tool, path = sys.argv[1], sys.argv[2]
txt = open(path, errors="replace").read()
ok_run = "sanitize_small ok" in txt
blocks = re.split(r"\n========= \n", txt)
listed = unlisted = 0
samples = []
for b in blocks:
    if "Race reported" in b:
        w = re.search(r"Write access at .*?(k_conv3x3_tc[24])<.*?\+0x(f{8}[0-9a-f]+)", b)
        reads = re.findall(r"and Read access at .*?(k_conv3x3_tc[24])<", b)
        if w and reads and "Race reported between Write access at" in b and b.count("Write access") == 1:
            listed += 1
        else:
            unlisted += 1
            samples.append(b[:600])
    elif re.search(r"Barrier error|Invalid __|Misaligned|out of bounds|Error:|Uninitialized|hazard", b) and "ERROR SUMMARY" not in b and "RACECHECK SUMMARY" not in b:
        unlisted += 1
        samples.append(b[:600])
summ = re.findall(r"(ERROR SUMMARY.*|RACECHECK SUMMARY.*)", txt)
print("%s: run %s; %d listed (tcgen05.alloc pair hand-off), %d UNLISTED; %s" % (
    tool, "completed" if ok_run else "DID NOT COMPLETE", listed, unlisted, "; ".join(summ)))
for smp in samples[:5]:
    print("--- unlisted ---\n" + smp)
sys.exit(0 if (ok_run and unlisted == 0) else 1)
