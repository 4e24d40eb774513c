"""ctypes prototypes for the later sections of include/dsdneo_b200.h (bound only if exported)."""
import ctypes as C


def _has(L, name):
    try:
        getattr(L, name)
        return True
    except AttributeError:
        return False


def bind(L):
    vp, sz, ci, cf = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    cd = C.c_double
    protos = {
        "dsdneo_b200_viterbi_k5_decode_batch": (ci, [vp, sz, ci, vp, ci, vp, sz, vp, ci, vp]),
        "dsdneo_b200_viterbi_k5_decode_batch_host": (ci, [vp, sz, ci, vp, ci, vp, sz, vp, ci]),
        "dsdneo_b200_nxdn_conv_decode_batch": (ci, [vp, vp, sz, ci, ci, vp, vp, sz, ci, vp]),
        "dsdneo_b200_nxdn_conv_decode_batch_host": (ci, [vp, vp, sz, ci, ci, vp, vp, sz, ci]),
        "dsdneo_b200_sym_class_from_synctype": (ci, [ci, ci, ci, vp]),
        "dsdneo_b200_symbolizer_create": (vp, [vp]),
        "dsdneo_b200_symbolizer_destroy": (None, [vp]),
        "dsdneo_b200_symbolizer_reset": (ci, [vp, vp]),
        "dsdneo_b200_symbolizer_set_class": (ci, [vp, vp]),
        "dsdneo_b200_symbolize_batch": (ci, [vp, vp, sz, ci, ci, ci, vp, vp]),
        "dsdneo_b200_fec_block_code_len": (ci, [ci]),
        "dsdneo_b200_fec_block_code_k": (ci, [ci]),
        "dsdneo_b200_fec_block_decode_batch": (ci, [ci, vp, vp, vp, ci, vp]),
        "dsdneo_b200_fec_block_decode_batch_host": (ci, [ci, vp, vp, vp, ci]),
        "dsdneo_b200_fec_golay_24_12_encode_batch": (ci, [vp, vp, ci, vp]),
        "dsdneo_b200_bptc_196x96_batch": (ci, [vp, ci, vp, vp, vp, ci, vp]),
        "dsdneo_b200_bptc_196x96_batch_host": (ci, [vp, ci, vp, vp, vp, ci]),
        "dsdneo_b200_bptc_128x77_batch": (ci, [vp, vp, vp, ci, vp]),
        "dsdneo_b200_bptc_128x77_batch_host": (ci, [vp, vp, vp, ci]),
        "dsdneo_b200_bptc_16x2_batch": (ci, [vp, vp, vp, ci, ci, vp]),
        "dsdneo_b200_bptc_16x2_batch_host": (ci, [vp, vp, vp, ci, ci]),
        "dsdneo_b200_p25_12_soft_llr_batch": (ci, [vp, vp, vp, ci, vp]),
        "dsdneo_b200_p25_12_soft_llr_batch_host": (ci, [vp, vp, vp, ci]),
        "dsdneo_b200_p25_12_soft_llr_list_batch": (ci, [vp, vp, vp, ci, ci, vp]),
        "dsdneo_b200_p25_12_soft_llr_list_batch_host": (ci, [vp, vp, vp, ci, ci]),
        "dsdneo_b200_p25_rs_decode_batch": (ci, [ci, vp, vp, vp, ci, vp]),
        "dsdneo_b200_p25_rs_decode_batch_host": (ci, [ci, vp, vp, vp, ci]),
        "dsdneo_b200_p25_rs_decode_erasures_batch": (ci, [ci, vp, vp, vp, ci, vp, vp, ci, vp]),
        "dsdneo_b200_p25_rs_decode_erasures_batch_host": (ci, [ci, vp, vp, vp, ci, vp, vp, ci]),
        "dsdneo_b200_p25_rs_soft_reliability_batch": (ci, [ci, vp, vp, vp, vp, ci, vp, ci, vp]),
        "dsdneo_b200_p25_rs_soft_reliability_batch_host": (ci, [ci, vp, vp, vp, vp, ci, vp, ci]),
        "dsdneo_b200_p25_word_decode_batch": (ci, [ci, vp, vp, vp, vp, ci, vp]),
        "dsdneo_b200_p25_word_decode_batch_host": (ci, [ci, vp, vp, vp, vp, ci]),
        "dsdneo_b200_bch_63_16_decode_batch": (ci, [vp, vp, vp, vp, ci, vp]),
        "dsdneo_b200_bch_63_16_decode_batch_host": (ci, [vp, vp, vp, vp, ci]),
        "dsdneo_b200_timing_enable": (ci, [ci]),
        "dsdneo_b200_timing_report": (ci, [C.c_char_p, sz]),
        "dsdneo_b200_frame_sync_create": (vp, [ci, vp, ci]),
        "dsdneo_b200_frame_sync_destroy": (None, [vp]),
        "dsdneo_b200_frame_sync_reset": (ci, [vp, vp]),
        "dsdneo_b200_frame_sync_search_batch": (ci, [vp, vp, sz, vp, vp, ci, vp, vp]),
        "dsdneo_b200_hb_cascade_create": (vp, [ci, ci, ci]),
        "dsdneo_b200_hb_cascade_destroy": (None, [vp]),
        "dsdneo_b200_hb_cascade_reset": (ci, [vp, vp]),
        "dsdneo_b200_hb_cascade_decim_batch": (ci, [vp, vp, sz, ci, ci, vp, sz, vp]),
        "dsdneo_b200_hb_cascade_decim_batch_host": (ci, [vp, vp, sz, ci, ci, vp, sz]),
        "dsdneo_b200_frontend_create": (vp, [vp]),
        "dsdneo_b200_frontend_destroy": (None, [vp]),
        "dsdneo_b200_frontend_reset": (ci, [vp, vp]),
        "dsdneo_b200_frontend_bank": (vp, [vp]),
        "dsdneo_b200_frontend_process": (ci, [vp, vp, sz, vp, sz, vp]),
        "dsdneo_b200_frontend_process_host": (ci, [vp, vp, sz, vp, sz]),
        "dsdneo_b200_frontend_submit_host": (C.c_longlong, [vp, vp, sz, vp, sz]),
        "dsdneo_b200_frontend_wait_host": (ci, [vp, C.c_longlong]),
        "dsdneo_b200_frontend_process_async": (ci, [vp, vp, sz, vp, sz, vp]),
        "dsdneo_b200_frontend_join": (ci, [vp, vp]),
        "dsdneo_b200_channelizer_design_prototype": (ci, [ci, ci, cd, C.POINTER(cf)]),
        "dsdneo_b200_channelizer_create": (vp, [ci, ci, ci, C.POINTER(cf)]),
        "dsdneo_b200_channelizer_destroy": (None, [vp]),
        "dsdneo_b200_channelizer_reset": (ci, [vp, vp]),
        "dsdneo_b200_channelizer_get_prototype": (ci, [vp, C.POINTER(cf), ci]),
        "dsdneo_b200_channelize": (ci, [vp, vp, sz, vp, sz, vp]),
        "dsdneo_b200_channelize_host": (ci, [vp, vp, sz, vp, sz]),
        "dsdneo_b200_stream_server_create": (vp, [sz, C.c_uint, ci, ci, ci]),
        "dsdneo_b200_stream_server_destroy": (None, [vp]),
        "dsdneo_b200_stream_server_make_current": (None, [vp]),
        "dsdneo_b200_stream_server_push": (sz, [vp, vp, sz, ci]),
        "dsdneo_b200_stream_server_close": (None, [vp]),
        "dsdneo_b200_stream_server_bump_generation": (None, [vp]),
        "dsdneo_b200_stream_server_set_power": (None, [vp, C.c_double]),
        "dsdneo_b200_stream_hook_read": (ci, [vp, vp, sz, vp]),
        "dsdneo_b200_stream_hook_return_pwr": (C.c_double, [vp]),
        "dsdneo_b200_stream_hook_output_rate_hz": (C.c_uint, []),
        "dsdneo_b200_stream_hook_output_kind": (ci, []),
        "dsdneo_b200_stream_hook_symbol_profile": (ci, [vp, vp, vp]),
        "dsdneo_b200_stream_hook_stream_generation": (C.c_uint32, []),
        "dsdneo_b200_symbol_capture_size": (sz, [sz, ci]),
        "dsdneo_b200_symbol_capture_pack": (ci, [vp, vp, vp, vp, sz, ci, vp]),
        "dsdneo_b200_symbol_capture_unpack": (C.c_longlong, [vp, sz, vp, vp, vp, vp, sz]),
        "dsdneo_b200_symbol_capture_write_file": (ci, [C.c_char_p, ci, vp, vp, vp, vp, sz]),
        "dsdneo_b200_mbe_synth_batch": (ci, [vp, vp, vp, ci, vp, vp, ci, vp]),
        "dsdneo_b200_mbe_synth_batch_host": (ci, [vp, vp, vp, ci, vp, vp, ci]),
        "dsdneo_b200_selftest_atan2f": (ci, [vp, vp, vp, ci, vp]),
        "dsdneo_b200_selftest_scale": (ci, [vp, vp, vp, ci, vp]),
    }
    for name, (res, args) in protos.items():
        if _has(L, name):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
