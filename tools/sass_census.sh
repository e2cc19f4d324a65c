#!/usr/bin/env bash
# Blackwell-native evidence (B200_PROFILING.md): per kernel family, how many tcgen05 / TMA / TMEM instructions the
# shipped SASS contains.  Runs without a GPU.  Usage: bash tools/sass_census.sh > profiles/r02_sass_census.md
set -euo pipefail
SO="$(cd "$(dirname "$0")/.." && pwd)/vocoder_b200/libfv_b200.so"
TMP="$(mktemp)"
cuobjdump -sass "$SO" > "$TMP"
echo "# SASS census of vocoder_b200/libfv_b200.so (cuobjdump -sass, sm_100a)"
echo
echo "Counts of SASS mnemonics per kernel family (all template instantiations summed): \`UTCHMMA\` = tcgen05.mma,"
echo "\`UTMALDG/UTMASTG/UTMAREDG\` = TMA tensor load / store / reduce, \`LDTM/STTM\` = tcgen05.ld / tcgen05.st, \`HMMA\` = legacy mma.sync."
echo
echo "| kernel family | instantiations | UTCHMMA | UTCHMMA.2CTA | UTMALDG | UTMASTG | UTMAREDG | LDTM | STTM | HMMA (legacy) |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"
python3 - "$TMP" <<'PY'
import re, sys, collections
fam = collections.defaultdict(collections.Counter)
keys = ("conv_tc_kernel", "mrf_fused_kernel", "snake_conv_kernel", "snake_aa", "dwconv_ln", "conv_post", "istft_ola",
        "pack_input", "conv_simt", "rowshift_probe", "umma_rate", "frame_audio", "spec_mag", "log_mel", "act_cast",
        "resample_linear", "noise_conv", "unpack_output")
cur = None
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((k for k in keys if k in m.group(1)), "other")
        fam[cur]["n"] += 1
        continue
    if cur is None:
        continue
    if re.search(r"\bUTCHMMA\.2CTA", line):
        fam[cur]["2cta"] += 1
        continue
    for pat, key in ((r"\bUTCHMMA", "umma"), (r"\bUTMALDG", "ldg"), (r"\bUTMASTG", "stg"), (r"\bUTMAREDG", "redg"),
                     (r"\bLDTM", "ldtm"), (r"\bSTTM", "sttm"), (r"\bHMMA", "hmma")):
        if re.search(pat, line):
            fam[cur][key] += 1
for k, c in sorted(fam.items(), key=lambda kv: -(kv[1]["umma"] + kv[1]["2cta"])):
    print(f"| `{k}` | {c['n']} | {c['umma']} | {c['2cta']} | {c['ldg']} | {c['stg']} | {c['redg']} | {c['ldtm']} | {c['sttm']} | {c['hmma']} |")
PY
rm -f "$TMP"
