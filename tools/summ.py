import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e); continue
    r=d["roofline"]; 
    print(f"{f}: value={d['value']:.0f} cyc/s ms/step={d['ms_per_step']:.2f} cycles/step={d['cycles_per_step']} cyc-only={d['cycles_only_vcycles_per_s']:.0f} e2e={d['e2e']['value']:.0f} jacL0={r['us_per_launch']:.1f}us frac={r['frac']:.3f} vfrac={r['vcycle_frac']:.3f} split={d['last_step_split_ms']}")
    pk=d["per_kernel_us"]; print("   ", {k:round(v['us'],1) for k,v in pk.items()})
