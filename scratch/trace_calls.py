"""Summarise QMC4_TRACE output: python scratch/trace_calls.py <stderr file> [rank]  — per qmc_step call duration, slow steps"""
import sys
rank = sys.argv[2] if len(sys.argv) > 2 else "0"
calls, cur = [], []
for line in open(sys.argv[1]):
    if 'qmc4 trace rank %s]' % rank not in line: continue
    if 'step:' in line:
        if cur: calls.append(cur)
        cur = []
        continue
    cur.append([float(x) for x in line.split(']')[1].replace('|', ' ').replace(':', ' ').split()])
if cur: calls.append(cur)
for i, c in enumerate(calls):
    per = c[-1][4] / len(c)
    print("call %d: %d steps, %.3f ms, %.4f ms/step, %d migration events" % (i, len(c), c[-1][4], per, sum(1 for r in c if r[-1] > 0)))
    # steps whose duration (move0 -> next move0) is > 2x the median
    d = [c[k + 1][1] - c[k][1] for k in range(len(c) - 1)]
    if not d: continue
    med = sorted(d)[len(d) // 2]
    for k, x in enumerate(d):
        if x > 2 * med:
            r = c[k]
            print("   slow step %d: %.3f ms (median %.3f): move %.3f scan-wait %.3f scan+est %.3f | host wait %.3f..%.3f exit %.3f mig %d" % (
                k, x, med, r[2] - r[1], 0.0, r[4] - r[2], r[7], r[8], r[9], int(r[-1])))
