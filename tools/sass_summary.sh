#!/bin/bash
# Static evidence from the built library (no GPU needed): which SASS instructions the hot kernels use.
#   bash tools/sass_summary.sh > profiles/r2b_sass_mnemonics.txt
lib=pose_splatter_b200/libpsplat.so
echo "# cuobjdump -sass $lib: per kernel, the number of instructions and of the mnemonics that carry the design"
echo "# LDGSTS = cp.async (the record ring), UBLKCP = cp.async.bulk (1-D TMA, behind PS_BWD_BULK), SYNCS = mbarrier,"
echo "# RED / REDG = red.global.add (gradient rows), FFMA2 = packed fp32 FMA (loss kernels), SHFL / VOTE / MATCH = warp exchange,"
echo "# DFMA = fp64 (parameter head), ATOMS = shared-memory atomics (bucket ranking), STG.E.128 / LDG.E.128 = 16-byte accesses"
cuobjdump -sass $lib | awk '
/Function :/ { if (name != "") flush(); name = $3; n = 0; delete c; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
    n++; op = $2; sub(/;$/, "", op);
    if (op ~ /^@/) { op = $3; sub(/;$/, "", op) }
    base = op; sub(/\..*/, "", base);
    if (base ~ /^(LDGSTS|UBLKCP|SYNCS|RED|REDG|ATOMG|ATOMS|FFMA2|FMUL2|FADD2|SHFL|VOTE|MATCH|DFMA|DMUL|DADD|MUFU|BAR|LDS|STS|FFMA|FMUL|FADD)$/) c[base]++;
    if (op ~ /^LDG\.E\.128/) c["LDG.128"]++;
    if (op ~ /^STG\.E\.128/) c["STG.128"]++;
}
function flush(   k, line) {
    line = "";
    for (k in c) line = line " " k "=" c[k];
    printf "%s\t%d instructions\t%s\n", name, n, line;
}
END { if (name != "") flush() }' | c++filt | sed -E 's/\(anonymous namespace\):://; s/\(PsGeometry[^\t]*//; s/^void //' | sort
