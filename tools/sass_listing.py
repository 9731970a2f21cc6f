#!/usr/bin/env python
"""tools/sass_listing.py — SASS listings of the shipped hot kernels + a table of what one pass of their main loop issues.

    python tools/sass_listing.py            (needs only cuobjdump; reads vit-adapter_b200/lib/obj/*.o)

Writes profiles/sass/<kernel>.txt (cuobjdump -sass of that one function, from the object that is linked into
lib/libmsda_b200.so) and profiles/sass/README.md with, per kernel: registers, static shared memory, resident CTAs per SM that
the register count allows, and the static instruction mix of the MAIN LOOP (the widest backward branch of the function =
the loop over queries / sorted points), i.e. instructions per loop pass.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, 'vit-adapter_b200', 'lib', 'obj')
OUT = os.path.join(ROOT, 'profiles', 'sass')

# (file stem, object, mangled-name regex, what it is)
KERNELS = [
    ('fwd_f32_L3', 'msda_fwd.o', r'msda_fwd_vec_kernelIfLi8ELi3ELi4ELi4ELi16ELb0E', 'forward fp32 D=32, Injector (3 levels x 4 points), 8 lanes x 4 ch'),
    ('fwd_f32_L1', 'msda_fwd.o', r'msda_fwd_vec_kernelIfLi8ELi1ELi4ELi4ELi16ELb0E', 'forward fp32 D=32, Extractor (1 level x 4 points)'),
    ('fwd_bf16_L3', 'msda_fwd_bf16.o', r'msda_fwd_vec_kernelI13__nv_bfloat16Li4ELi3ELi4ELi4ELi16ELb0E', 'forward bf16 D=32, Injector, 4 lanes x 8 ch'),
    ('fwd_bf16_L1', 'msda_fwd_bf16.o', r'msda_fwd_vec_kernelI13__nv_bfloat16Li4ELi1ELi4ELi4ELi16ELb0E', 'forward bf16 D=32, Extractor'),
    ('bwd_f32_L3', 'msda_bwd.o', r'msda_bwd_vec_kernelIfLi8ELi3ELi4ELi3ELb0E', 'backward fp32 D=32, Injector, 8 lanes x 4 ch'),
    ('bwd_f32_L1', 'msda_bwd.o', r'msda_bwd_vec_kernelIfLi8ELi1ELi4ELi3ELb0E', 'backward fp32 D=32, Extractor'),
    ('bwd_bf16_L3', 'msda_bwd_bf16.o', r'msda_bwd_vec_kernelI13__nv_bfloat16Li8ELi3ELi4ELi3ELb0E', 'backward bf16 D=32, Injector (fp32 accumulator)'),
    ('bwd_bf16_L1', 'msda_bwd_bf16.o', r'msda_bwd_vec_kernelI13__nv_bfloat16Li8ELi1ELi4ELi3ELb0E', 'backward bf16 D=32, Extractor'),
    ('fwd_fused_f32_L3', 'msda_fwd.o', r'msda_fwd_vec_kernelIfLi8ELi3ELi4ELi4ELi16ELb1E', 'fused forward fp32 (softmax + locations in registers), Injector'),
    ('bwd_fused_f32_L3', 'msda_bwd.o', r'msda_bwd_vec_kernelIfLi8ELi3ELi4ELi3ELb1E', 'fused backward fp32, Injector'),
    ('bwd_cell_f32_L1', 'msda_bwd_cell.o', r'msda_bwd_cell_kernelIfLi4ELi1ELi4ELi3E', 'cell-bucketed backward fp32 D=32, Extractor (opt-in)'),
    ('bwd_cell_f32_L3', 'msda_bwd_cell.o', r'msda_bwd_cell_kernelIfLi4ELi3ELi4ELi3E', 'cell-bucketed backward fp32 D=32, Injector (opt-in)'),
    ('bwd_sorted_walk_f32_L1', 'msda_bwd_sorted.o', r'msda_bwd_sorted_kernelIfLi8ELi1ELi4ELi3E', 'slab-sorted backward, walker fp32 D=32, Extractor (main loop = one batch of 32 samples)'),
    ('bwd_sorted_walk_bf16_d64_L1', 'msda_bwd_sorted_bf16.o', r'msda_bwd_sorted_kernelI13__nv_bfloat16Li16ELi1ELi4ELi3E', 'slab-sorted backward, walker bf16 D=64, Extractor (north-star 16 x 64)'),
    ('bwd_sorted_hist_L1', 'msda_bwd_sorted.o', r'msda_sort_part_kernelILi1ELi4ELb0E', 'slab-sorted backward, histogram pass of the counting sort'),
    ('bwd_sorted_prefix', 'msda_bwd_sorted.o', r'msda_sort_prefix_kernel', 'slab-sorted backward, per-key prefix over the parts of the counting sort'),
    ('bwd_sorted_scatter_L1', 'msda_bwd_sorted.o', r'msda_sort_part_kernelILi1ELi4ELb1E', 'slab-sorted backward, scatter pass of the counting sort (key scan + cursors, then one ATOMS per sample)'),
]
MNEMONICS = ['LDG.E.128', 'LDG.E.64', 'LDG.E ', 'LDGSTS', 'LDS', 'STS', 'REDG.E.ADD.F32x4', 'ATOMS', 'SHFL', 'FFMA2', 'FMUL2', 'FFMA ', 'FMUL ', 'FADD', 'IMAD.WIDE',
             'BAR.SYNC', 'STG']


def sh(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout


def functions(obj):
    out = sh(['cuobjdump', '-sass', obj])
    return re.findall(r'Function : (\S+)', out)


def res_usage(obj):
    out = sh(['cuobjdump', '-res-usage', obj])
    res = {}
    for m in re.finditer(r'Function (\S+):\s*\n\s*(.*)', out):
        res[m.group(1)] = m.group(2)
    return res


def main_loop(lines):
    """(first, last) indices of the widest backward branch."""
    addr = {}
    ins = []
    for l in lines:
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
        if m:
            addr[int(m.group(1), 16)] = len(ins)
            ins.append((int(m.group(1), 16), m.group(2)))
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r'\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr and (best is None or a - tgt > best[2]):
                best = (addr[tgt], i, a - tgt)
    return ins, best


def main():
    os.makedirs(OUT, exist_ok=True)
    rows = []
    for stem, obj, pat, what in KERNELS:
        path = os.path.join(OBJ, obj)
        if not os.path.exists(path):
            sys.exit('missing %s: build the library first' % path)
        names = [n for n in functions(path) if re.search(pat, n)]
        if not names:
            print('no function matches', pat)
            continue
        name = names[0]
        sass = sh(['cuobjdump', '-sass', '-fun', name, path])
        with open(os.path.join(OUT, stem + '.txt'), 'w') as f:
            f.write('// %s\n// %s\n// cuobjdump -sass -fun %s vit-adapter_b200/lib/obj/%s\n' % (what, sh(['cu++filt', name]).strip(), name, obj))
            # keep address + instruction; drop the 128-bit encodings (two hex words per instruction, 2/3 of the bytes)
            for line in sass.splitlines():
                line = re.sub(r'\s*/\* 0x[0-9a-f]{16} \*/\s*$', '', line)
                if line.strip():
                    f.write(line.rstrip() + '\n')
        ins, loop = main_loop(sass.splitlines())
        ru = res_usage(path).get(name, '')
        reg = int(re.search(r'REG:(\d+)', ru).group(1)) if re.search(r'REG:(\d+)', ru) else 0
        smem = int(re.search(r'SHARED:(\d+)', ru).group(1)) if re.search(r'SHARED:(\d+)', ru) else 0
        body = ins[loop[0]:loop[1] + 1] if loop else ins
        counts = {m: sum(1 for _, t in body if re.search(r'(^|\s)' + re.escape(m.strip()) + (r'(\s|\.|$)' if m.endswith(' ') else ''), t)) for m in MNEMONICS}
        ctas = min(32, 65536 // (max(reg, 1) * 256)) if reg else 0
        rows.append((stem, what, reg, smem, ctas, len(ins), len(body), counts))
    with open(os.path.join(OUT, 'README.md'), 'w') as f:
        f.write('# SASS listings of the shipped kernels (sm_100a, nvcc 12.9)\n\n')
        f.write('Generated by `python tools/sass_listing.py` from `vit-adapter_b200/lib/obj/*.o`, the objects linked into '
                '`lib/libmsda_b200.so`. One file per kernel (`cuobjdump -sass -fun <name>`).\n\n'
                '"main loop" = the widest backward branch of the function: the loop over a warp\'s queries (vector kernels; one '
                'pass = 32/G queries x all L*P points, fully unrolled) or over the batches of 32 sorted points (cell kernel). '
                'Counts are STATIC instructions inside that loop, i.e. what one pass issues when every branch is taken once; '
                'resident CTAs/SM is what the register count allows at 256 threads per CTA (the launch bounds ask for 3 or 4).\n\n')
        f.write('| kernel | what | regs | static smem B | CTAs/SM by regs | SASS instr | main loop instr | ' + ' | '.join(m.strip() for m in MNEMONICS) + ' |\n')
        f.write('|---|---|--:|--:|--:|--:|--:|' + '--:|' * len(MNEMONICS) + '\n')
        for stem, what, reg, smem, ctas, n, nb, c in rows:
            f.write('| `%s` | %s | %d | %d | %d | %d | %d | ' % (stem, what, reg, smem, ctas, n, nb) + ' | '.join(str(c[m]) for m in MNEMONICS) + ' |\n')
        f.write('\nThese are WARP instructions. In the vector kernels one pass serves 32/G queries at once (G lanes per query), so one '
                '`LDG.E.128` gathers one corner row for each of them: `fwd_f32_L1` issues 16 `LDG.E.128` per pass = 4 points x 4 corners '
                'for 4 queries (16 sampled points: 20 instructions per point), `bwd_f32_L1` 16 `LDG.E.128` + 16 `REDG.E.ADD.F32x4` '
                '(4 + 4 rows per point; 46 instructions per point - the same figure ncu reports as inst_executed / points). '
                'Per pass the kernels serve `(32/G) * L*P` points: 48 (Injector, G=8), 16 (Extractor, G=8), 96 / 32 (bf16 forward, G=4).\n')
    print(open(os.path.join(OUT, 'README.md')).read())


if __name__ == '__main__':
    main()
