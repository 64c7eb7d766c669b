"""Static SASS instruction counts of one sub-step kernel instantiation, per phase and per source line -- no GPU needed.

    python tools/sass_phases.py [MODE] [path/to/libabm_b200.so] [--lines N]

MODE is the template argument of abm_sweep_kernel (259 night, 135 go-out, 451 noon, 75 go-home, 1 flush; default
259).  The library must have been built with -lineinfo (the default build is).  Phases are the same as in
tools/ncu_phases.py: every `__device__` function and every `// ====... <title>` banner of the kernel body.

Static counts are not executed counts (loops, divergent funnels), but the hot path of this kernel is almost all
straight-line predicated code, so the per-phase static counts track the `smsp__inst_executed` shares of
profiles/r01_ncu_phases_final_night.txt closely enough to screen an "instruction diet" change before spending GPU time
on it; spill traffic (STL / LDL) and the opcode mix are printed as well.
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'covid19-agent-based-model_b200', 'csrc', 'abm_kernels.cuh')
CUDA_BIN = os.environ.get('CUDA_BIN', '/usr/local/cuda/bin')


def phase_marks(src):
    marks, in_kernel = [], False
    for no, line in enumerate(open(src), 1):
        m = re.match(r'\s*(?:template\s*<[^>]*>\s*)?__device__\s+(?:__forceinline__\s+)?[\w:<> ]+?\s+(\w+)\s*\(', line)
        if m:
            marks.append((no, 'fn ' + m.group(1)))
            in_kernel = False
        if re.match(r'abm_sweep_kernel\(', line):
            marks.append((no, 'kernel prologue'))
            in_kernel = True
        b = re.match(r'\s*// =+ (.+)$', line)
        if b and in_kernel:
            marks.append((no, b.group(1).strip()[:40]))
        if re.match(r'__global__', line) and not in_kernel:
            marks.append((no, 'other kernels'))
    return sorted(marks)


def main():
    argv = list(sys.argv[1:])
    if '--lines' in argv:
        j = argv.index('--lines'); top_arg = argv[j + 1]; del argv[j:j + 2]
    else:
        top_arg = '12'
    args = [a for a in argv if not a.startswith('--')]
    top = int(top_arg)
    mode = args[0] if args else '259'
    lib = args[1] if len(args) > 1 else os.path.join(ROOT, 'covid19-agent-based-model_b200', 'libabm_b200.so')
    tmp = tempfile.mkdtemp(prefix='sass_')
    subprocess.run([os.path.join(CUDA_BIN, 'cuobjdump'), '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, check=True,
                   capture_output=True)
    cubin = sorted(f for f in os.listdir(tmp) if f.endswith('.cubin'))[0]
    text = subprocess.run([os.path.join(CUDA_BIN, 'nvdisasm'), '-g', '-c', os.path.join(tmp, cubin)], check=True,
                          capture_output=True, text=True).stdout
    want = f'abm_sweep_kernelILj{mode}ELb0E'        # the instantiation without the timeline stamps
    marks = phase_marks(SRC)

    def phase_of(line_no):
        name = 'other'
        for first, n in marks:
            if first <= line_no:
                name = n
            else:
                break
        return name

    inside, cur_file, cur_line = False, None, 0
    per_phase, per_line, ops = Counter(), Counter(), Counter()
    outside_src = Counter()
    for row in text.splitlines():
        if row.startswith('\t.section') or row.startswith('//-----'):
            inside = want in row and '.text.' in row
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', row)
        if m:
            cur_file, cur_line = m.group(1), int(m.group(2))
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', row)
        if not m:
            continue
        op = m.group(1)
        ops[op.split('.')[0]] += 1
        if cur_file and cur_file.endswith('abm_kernels.cuh'):
            per_phase[phase_of(cur_line)] += 1
            per_line[cur_line] += 1
        else:                      # intrinsics headers (__ldcs, __shfl_sync ...): charged to the line that was current before
            outside_src[os.path.basename(cur_file or '?')] += 1
            per_phase['(cuda headers: loads, shuffles, votes)'] += 1
    total = sum(per_phase.values())
    if not total:
        sys.exit(f'no SASS found for {want} in {lib}')
    print(f'abm_sweep_kernel<{mode}>: {total} static SASS instructions')
    for name, n in per_phase.most_common():
        print(f'  {name:44s} {n:6d}  {100 * n / total:5.1f}%')
    print('opcode mix:', ', '.join(f'{k} {v}' for k, v in ops.most_common(14)))
    print('local-memory traffic (spills): STL', ops.get('STL', 0), ' LDL', ops.get('LDL', 0))
    src_lines = open(SRC).read().splitlines()
    print(f'top {top} source lines:')
    for line_no, n in per_line.most_common(top):
        print(f'  {line_no:5d} {n:5d}  | {src_lines[line_no - 1].strip()[:110]}')


if __name__ == '__main__':
    main()
