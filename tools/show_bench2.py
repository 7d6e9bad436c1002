#!/usr/bin/env python
"""Compact view of a bench.py JSON line (round-2 layout)."""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"== {path}")
    print(f"N={d['n_gpus']} value {d['value']:.0f} ({d['ms_per_step']:.1f} ms)  e2e {d['e2e']['value']:.0f}  pageable {d['e2e_pageable']['value']:.0f} ({d['e2e_pageable']['fraction_of_pinned']:.3f})")
    oc = d.get('e2e_one_call') or {}
    print("  one_call:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in oc.items() if k not in ('timing', 'note')})
    print("  weak:", round(d['weak_scaling']['value']), " int:", round(d['integer_valued_inputs']['value']), " other mode:", round(d['other_superposition_mode']['value']))
    print("  roofline:", d['roofline']['kernel'], round(d['roofline']['frac'], 3), "util", round(d['roofline']['fp64_pipe_util'], 3), "| other", d['roofline_other_fp64']['kernel'], round(d['roofline_other_fp64']['fp64_pipe_util'], 3), "| pipeline", round(d['pipeline_fp64']['fp64_pipe_util'], 3))
    c3 = d.get('config3')
    if c3:
        for b in c3['batches']:
            print(f"  config3 {b['total_spectra']}: value {b['value']:.0f} e2e {b['e2e']:.0f} ({b['e2e_fraction_of_value']:.3f}) pageable {b['e2e_pageable']:.0f} ({b['e2e_pageable_fraction_of_pinned']:.3f})")
        print("  config3 pipeline util", round(c3['pipeline_fp64']['fp64_pipe_util'], 3), "fit_iter util", round(c3['roofline_fit_iter']['fp64_pipe_util'], 3), "serial ms", {k: round(v, 1) for k, v in c3['kernel_ms_serial_step'].items()})
    s = d.get('superposition_vec')
    if s:
        print(f"  config4: {s['evals_per_s']:.3e} evals/s ({s['ms']:.1f} ms) e2e {s['e2e']['evals_per_s']:.3e} ({s['e2e']['fraction_of_device_resident']:.3f}) one_call {s.get('one_call')}")
    print("  cpu:", d.get('cpu_baseline'), "\n  parity:", d.get('parity_sample') and d['parity_sample']['pass'], d.get('parity_per_rank'))
    print("  kernel ms:", {k: round(v, 1) for k, v in d['kernel_ms_serial_step'].items()}, "wall", round(d['bench_wall_s']))
