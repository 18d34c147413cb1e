summ() {
python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        def one(tag,d):
            print(tag, round(d['value']/1e9,4),'G/s', round(d['ms_per_step'],1),'ms', {k:round(v,1) for k,v in d['pipeline']['kernel_ms_per_step'].items()},
                  'e2e',round(d['e2e']['value']/1e9,4), 'roof',round(d['roofline']['frac'],3), 'mapped',d['mapped_reads'],'truth',d['truth_concordant_reads'])
            c=d['pipeline']['counters_per_step']; print('   ', {k:c[k] for k in ('steps','capped_queries','overflow_queries','part_sort_steps','seg_sort_steps','sync_points')})
            print('    conc', d.get('concordance'))
        one('c2' if d['config']['baseline_config']=='configs[1]' else 'c3', d)
        if d.get('latency'): print('    latency p50', round(d['latency']['p50'],3), 'p90', round(d['latency']['p90'],3))
        if d.get('cpu_baseline'): print('    cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
        if d.get('config3'): one('c3@1', d['config3'])
PY
}
