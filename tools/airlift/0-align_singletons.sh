#!/bin/bash
# Drop-in for AirLift's src/0-align_singletons.sh (reference line 12): single-end re-alignment of the reads whose mate was not
# extracted, with the B200 build of the minimap2 fork.  Same positional arguments and outputs as the reference script.
BINDIR=$1
REF=$2
FASTQ=$3
OUT_PREFIX=$4
THREADS=$5
THREAD_SORT=$6
SAMPLE=$7
MAXMEM=$8

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
MM2="${MINIMAP2_B200:-${HERE}/../../build/minimap2-b200}"
SAMTOOLS="${SAMTOOLS:-samtools}"
TIME=()
[ -x /usr/bin/time ] && TIME=(/usr/bin/time -v -p -o "${OUT_PREFIX}.time")

"${TIME[@]}" "${MM2}" -ax sr --gpus "${GPUS:-1}" -R "@RG\tID:${SAMPLE}\tSM:${SAMPLE}\tPL:illumina\tLB:${SAMPLE}" -t "${THREADS}" "${REF}" "${FASTQ}" \
	| "${SAMTOOLS}" view -h -F4 | "${SAMTOOLS}" sort -l5 -@ "${THREAD_SORT}" -m "${MAXMEM}" > "${OUT_PREFIX}.bam"
