"""Opcode histogram of the fused step kernel's SASS (cuobjdump -sass on the built object): the evidence that the data path is
the Blackwell one (TMA bulk copies UBLKCP, mbarrier SYNCS, bulk-group waits) and that the hot path stays out of local memory.
    python profiles/sass_summary.py > profiles/sass_summary.txt        (needs only the built objects, no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "robotic-warehouse_b200", "build", "rw_inst_0.o")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
want = {"_Z14rw_main_kernelILi0E8RwStaticILi11ELi10ELi32ELi1ELi0ELi4ELi4EELi896ELb0EE": "STEP  tiny-4ag (full static, 896 threads)",
        "_Z14rw_main_kernelILi1E8RwStaticILi11ELi10ELi32ELi1ELi0ELi4ELi4EELi896ELb0EE": "OBS   tiny-4ag",
        "_Z14rw_main_kernelILi2E8RwStaticILi11ELi10ELi32ELi1ELi0ELi4ELi4EELi896ELb0EE": "RESET tiny-4ag",
        "_Z14rw_main_kernelILi0E8RwStaticILi11ELi10ELi32ELi1ELi0ELi4ELi4EELi896ELb1EE": "STEP  tiny-4ag, bit form of the rows (host transport)"}
groups = [("TMA bulk copies (cp.async.bulk)", r"^UBLKCP"), ("TMA command flush / bulk-group wait", r"^UTMACMDFLUSH|^UTMACCTL|^DEPBAR"),
          ("mbarrier (SYNCS.*)", r"^SYNCS"), ("named / CTA barriers (BAR)", r"^BAR"), ("proxy fences (FENCE)", r"^FENCE"),
          ("shared loads (LDS)", r"^LDS"), ("shared stores (STS)", r"^STS"), ("global loads (LDG)", r"^LDG"), ("global stores (STG)", r"^STG"),
          ("generic loads/stores (LD/ST)", r"^(LD|ST)(\.|$)"), ("local memory (LDL/STL)", r"^(LDL|STL)"), ("atomics / reductions", r"^(ATOM|RED|ATOMS|ATOMG)"),
          ("warp votes / shuffles (VOTE/SHFL/MATCH)", r"^(VOTE|SHFL|MATCH)"), ("programmatic dependent launch (ACQBULK/griddep)", r"^(ACQBULK|CCTL)"),
          ("tensor core (HMMA/UTC*MMA/…)", r"^(HMMA|IMMA|DMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCOMMA|UTCMMA)")]
print("# SASS opcode summary of rw_main_kernel (cuobjdump -sass robotic-warehouse_b200/build/rw_inst_0.o; nvcc 12.9, sm_100a)")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    key = next((k for k in want if name.startswith(k)), None)
    if not key:
        continue
    ops = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)]
    hist = collections.Counter(ops)
    print(f"\n## {want[key]}\n   {name}\n   {len(ops)} instructions")
    for label, pat in groups:
        sel = {o: c for o, c in hist.items() if re.match(pat, o)}
        tot = sum(sel.values())
        top = ", ".join(f"{o} x{c}" for o, c in sorted(sel.items(), key=lambda kv: -kv[1])[:6])
        print(f"   {label:52s} {tot:5d}   {top}")
