#!/bin/bash
# Drop-in for AirLift's src/0-align_reads.sh (reference: `bwa mem -R ... | samtools view -h -F4 | samtools sort`, line 13):
# the paired re-alignment of reads from updated / retired regions against the NEW reference, done by the B200 build of
# AirLift's own minimap2 fork (`-ax sr`).  Same positional arguments, same outputs (${OUT_PREFIX}.bam, ${OUT_PREFIX}.time),
# so run_pipeline.sh:110 and the merge steps after it (run_pipeline.sh:116-119) need no change.
#   MINIMAP2_B200  the binary (default: build/minimap2-b200 of this repository)     GPUS  number of GPUs to shard reads over
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

set -x
"${TIME[@]}" "${MM2}" -ax sr --gpus "${GPUS:-1}" -R "@RG\tID:${SAMPLE}\tSM:${SAMPLE}\tPL:illumina\tLB:${SAMPLE}" -t "${THREADS}" "${REF}" "${FASTQ}_1.fastq" "${FASTQ}_2.fastq" \
	| "${SAMTOOLS}" view -h -F4 | "${SAMTOOLS}" sort -l5 -@ "${THREAD_SORT}" -m "${MAXMEM}" > "${OUT_PREFIX}.bam"
