#!/bin/bash
# what the GPU box offers the output sink: cores, memory, tmpfs room, NUMA layout
echo "nproc: $(nproc)"; free -g | head -2; df -h /dev/shm /tmp 2>/dev/null; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ; nvidia-smi topo -m 2>/dev/null | head -12
