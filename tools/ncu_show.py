import csv, sys
for tag in sys.argv[1:]:
    rows = [r for r in csv.reader(open('gpurun_out/ncu_%s.csv' % tag)) if len(r) > 10]
    hdr = rows[0]
    out = []
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d['ID'] == '0':
            n = d['Metric Name'].replace('smsp__average_warps_issue_stalled_', 'st_').replace('_per_issue_active.ratio', '')
            out.append("%s=%s" % (n.replace('.avg.pct_of_peak_sustained_active', '%').replace('.sum', ''), d['Metric Value']))
    print(tag, " ".join(out))
