"""fallingsandsurvival_b200 -- B200-native core for FallingSandSurvival's World::tick / World::tickTemperature.

csrc/            CUDA kernels (sm_100a) and the C ABI of include/fss.h  -> libfss_b200.so
world.py         host-side mirror of the reference's `World` for this path (ctypes over the C ABI)
strips.py        row-strip sharding plan and schedule for multi-GPU worlds
materials.py     Materials::MATERIALS and the Tiles::create* colour recipes as data
worlds.py        synthetic benchmark / test worlds
abi.py           ctypes / numpy mirrors of the structs in include/fss.h
"""
