#!/usr/bin/env bash
# Builds libb200attn_trace.so (-DB200_TC_TRACE -DB200_DEC_TRACE) next to the product library, then rebuilds the product library.
set -e
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"; c="$here/../scratchpad_b200/csrc"
B200_TC_TRACE=1 B200_DEC_TRACE=1 bash "$c/build.sh" >/dev/null && cp "$c/libb200attn.so" "$c/libb200attn_trace.so"
bash "$c/build.sh"
