#!/usr/bin/env python
"""Print the headline fields of a bench.py JSON line (file argument or stdin)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
print("value", round(d["value"], 1), d["unit"], "| e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
      "| launches", d.get("gpu_launches"), "| fp64 pipe", d.get("pipeline_fp64", {}).get("fp64_pipe_util"),
      "| roofline", d["roofline"].get("kernel"), d["roofline"].get("frac_of_fp64_pipe", d["roofline"].get("frac")),
      "| clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
