#!/bin/sh
# Evidence that the kernels are Blackwell-native: SASS / PTX mnemonic counts of the built library.
#   tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=${1:-safe_denoiser_b200/libsdn_repel.so}
echo "# cuobjdump -sass $LIB | grep -c <mnemonic>   (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP)"
cuobjdump -sass "$LIB" > /tmp/sdn_sass.txt
for m in UTCHMMA UTCBAR LDTM STTM UTMALDG UBLKCP UTMAPF "SYNCS" "HMMA\." "HGMMA" "LDGSTS" "REDG" "ATOMG"; do
  printf "%-10s %s\n" "$m" "$(grep -c "$m" /tmp/sdn_sass.txt)"
done
echo "# per kernel: tcgen05.mma (UTCHMMA) / TMEM loads (LDTM) / TMEM stores (STTM) / TMA tile loads (UTMALDG) / bulk copies (UBLKCP)"
awk '/Function : /{name=$3} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /STTM/{c[name]++} /UTMALDG/{d[name]++} /UBLKCP/{e[name]++} END{for(k in a) printf "%s  mma=%d ldtm=%d sttm=%d utmaldg=%d ublkcp=%d\n", k, a[k], b[k], c[k], d[k], e[k]}' /tmp/sdn_sass.txt | c++filt | sort
echo "# inline PTX in safe_denoiser_b200/csrc (the library carries SASS only: -gencode arch=compute_100a,code=sm_100a): grep -c per file"
for m in "tcgen05.mma" "tcgen05.ld" "tcgen05.st" "tcgen05.commit" "tcgen05.alloc" "cp.async.bulk.tensor" "cp.async.bulk.shared" "mbarrier.try_wait" "st.async" "createpolicy" "fence.proxy.async"; do
  printf "%-24s %s\n" "$m" "$(grep -c "$m" safe_denoiser_b200/csrc/*.cu safe_denoiser_b200/csrc/*.cuh | grep -v ':0' | sed 's#safe_denoiser_b200/csrc/##' | tr '\n' ' ')"
done
