#!/bin/bash
# profiles/r02_sass_{tc,tc_x3,halo,dcn_persist,decode,track}.txt + a mnemonic census: cuobjdump -sass of the built library.
# usage: tools/sass_dump.sh   (no GPU needed)
set -e
cd "$(dirname "$0")/.."
SO=centertrack_b200/libctb200.so
dump() { cuobjdump -sass -fun "$1" $SO 2>/dev/null | sed -e 's#/\* 0x[0-9a-f]* \*/##' -e 's/[ \t]*$//' > "$2"; }
dump _ZN3ctb14conv_tc_kernelILb0EEEvNS_6TcArgsE14CUtensorMap_st profiles/r02_sass_tc.txt
dump _ZN3ctb14conv_tc_kernelILb1EEEvNS_6TcArgsE14CUtensorMap_st profiles/r02_sass_tc_x3.txt
dump _ZN3ctb16conv_halo_kernelENS_8HaloArgsE14CUtensorMap_st profiles/r02_sass_halo.txt
dump _ZN3ctb18dcn_persist_kernelENS_6TcArgsEi14CUtensorMap_st profiles/r02_sass_dcn_persist.txt
dump _ZN3ctb13decode_kernelILb1EEEvNS_10DecodeArgsE profiles/r02_sass_decode.txt
dump _ZN3ctb17track_step_kernelENS_9TrackArgsE profiles/r02_sass_track.txt
{
  echo "# Blackwell-native mnemonics per kernel (cuobjdump -sass of $SO; tools/sass_dump.sh)"
  for f in tc tc_x3 halo dcn_persist decode track; do
    echo "## r02_sass_$f.txt ($(grep -c '^ *\/\*[0-9a-f]*\*\/' profiles/r02_sass_$f.txt) instructions)"
    for m in UTCHMMA UTCBAR LDTM STTM UTMALDG UBLKCP UTCATOMSWS SYNCS ACQBULK PREEXIT HFMA2.BF16 HMUL2.BF16 FFMA2 LDS.128 LDG.E.128 ' HMMA'; do
      n=$(grep -c "$m" profiles/r02_sass_$f.txt || true)
      [ "$n" != "0" ] && echo "  $m: $n"
    done
  done
} > profiles/r02_sass_census.txt
cat profiles/r02_sass_census.txt
