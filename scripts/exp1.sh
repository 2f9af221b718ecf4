python scripts/time_encoder.py 2>&1 | tail -12
echo "=== HALO debug0"; python scripts/time_conv.py
echo "=== HALO debug1 (no MMA)"; B200_CONV_DEBUG=1 python scripts/time_conv.py
echo "=== HALO debug2 (no TMA)"; B200_CONV_DEBUG=2 python scripts/time_conv.py
echo "=== PLAIN debug1 (no MMA)"; B200_CONV_NO_HALO=1 B200_CONV_DEBUG=1 python scripts/time_conv.py
echo "=== PLAIN debug2 (no TMA)"; B200_CONV_NO_HALO=1 B200_CONV_DEBUG=2 python scripts/time_conv.py
