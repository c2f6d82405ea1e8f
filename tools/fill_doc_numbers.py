"""Fills the {PLACEHOLDER} numbers of DESIGN.md from a bench.py default line: python tools/fill_doc_numbers.py profiles/r02_bench_final.json"""
import json, re, sys
d = json.load(open(sys.argv[1]))
ex = {e["metric"].split("cfg")[1][0]: e for e in d["extra_configs"] if "metric" in e}
vals = {
    "CFG3V": f"{d['value']:.0f}", "CFG3E": f"{d['e2e']['value']:.0f}", "CFG3S": f"{d['e2e']['synchronous_call']['value']:.0f}",
    "FRAC": f"{d['roofline']['frac']:.2f}", "CONVMS": f"{d['roofline']['kernel_ms_per_step']:.2f}", "PARSE": f"{d['breakdown']['parse_ms_per_step']:.3f}",
    "CFG2V": f"{ex['2']['value']:.0f}", "CFG4V": f"{ex['4']['value']:.0f}", "CFG5V": f"{ex['5']['value']:.0f}",
    "TF32V": f"{d['tf32']['value']:.0f}", "TF32F": f"{d['tf32']['roofline']['frac']:.2f}",
}
for path in sys.argv[2:] or ["DESIGN.md"]:
    s = open(path).read()
    for k, v in vals.items():
        s = s.replace("{" + k + "}", v)
    left = re.findall(r"\{[A-Z0-9]+\}", s)
    open(path, "w").write(s)
    print(path, "filled;", "left:", left)
