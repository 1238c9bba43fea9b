"""CPU oracle for the MKID SDR hot path -- TEST INFRASTRUCTURE ONLY.

Every module here is a plain NumPy / pure-Python / C restatement of the
reference's algorithm (creanero/MKIDS_SDR), each function citing the
reference file:line it follows.  Nothing under ``mkids_sdr_b200/`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / reported CPU baseline, never as the product path.

Pinned against the reference's own fixtures (see tests/golden/make_golden.py):
  * lut.py        <- ChannelizerControls/dac.npy.npz  (bit-exact)
  * fixed.py      <- castBin call sites in lib/set_alpha.py, set_base_thresh.py,
                     set_svf.py (known answers), Utils/bin.py run under py3 for
                     the py3-safe functions
  * control.py    <- LUT/*.txt FIR tap files (quantised taps), ch_snap_0.txt
  * trigger.py    <- ch_snap_0.txt
  * decode.py, packetmaster_core.c
                  <- oracle/_ref/libpm_ref_*.so: the reference's OWN receive loop
                     (PacketMaster.c:304-397), extracted as text at build time and
                     compiled by build_pm_ref.py with the process shell stubbed
                     (pm_ref.py is its ctypes front; tests/test_pm_ref_cpu.py);
                     the bitfield layout is pinned by ROACH_Pulses.py:805-811.
and against OUTPUTS OF THE REFERENCE'S OWN CODE executed in the dev container (the method sources are read from the
reference tree at run time, Python-2 -> 3 edits applied, run against a recording roach / mock widgets; only the
numerical outputs are stored: tests/golden/make_golden_refrun.py -> refrun_golden.npz,
make_golden_analysis.py -> analysis_golden.npz; tests/test_oracle_golden.py::*reference_run*):
  * lut.py        <- ROACH_Setup_DAC.py define_DAC_LUT / freqCombLUT / define_DDS_LUT / select_bins / write_LUTs
                     (12 tones, seed-1000 phases, DDS phases, DRAM image: bit-identical)
  * control.py    <- ROACH_Pulses.py loadFIRcoeffs, loadIQcenters, loadThresholds (bit-identical)
  * decode.py     <- ROACH_Pulses.py readPulses (10 steps, ring wraps: bit-identical)
  * trigger.py    <- the trigger loops of pulse_triggering_v2.py and pulse_triggering.py (hit lists identical)
  * fixed.py      <- Utils/bin.py, Utils/binTools.py run with Python-2 division (all functions, bit-identical)
  * template.py   <- lib/pulses.py MakeTemplate (bit-identical)
  * spectra.py    <- ArconsDashboard.py image_Worker methods (bit-identical)
  * channelizer.py: PARITY UNPINNED -- the firmware data plane is absent from
                     the reference; this float64 model is the parity definition.
"""
