"""sdr_channelizer_b200 — B200-native (sm_100a) polyphase channelizer + channelized PDW extraction.

Drop-in for the one hot path of cwozny/sdr_channelizer (matlab/channelizer_example.m,
matlab/create_pdws_channelized.m on recordings in the cpp/IqPacket.h format).  The product is the C
ABI in include/channelizer.h (libchannelizer.so); this package is its Python host-side mirror.
There is no CPU path: without the built library and a B200 the compute calls raise.
"""
from ._lib import (CHZ_OPT_CHUNK_ROWS, CHZ_OPT_FORCE_PATH, CHZ_OPT_PDW_EVENT_PATH, CHZ_OPT_RETAIN, LIB_PATH, ChannelizerError, IqInfo, Pdw,
                   PdwParams, lib)
from .channelizer import (Channelizer, IqRecording, PdwTable, channelizer_example, create_pdws, create_pdws_channelized, design_prototype, event_peak_time,
                          next_event_time, predict_event, read_iq, spectrogram_my_iq, stft, unpack_ptr, write_iq)
from .sharding import plan_time_shards, stitch_rows

__all__ = ["Channelizer", "PdwTable", "channelizer_example", "IqRecording", "read_iq", "write_iq", "design_prototype", "unpack_ptr",
           "create_pdws_channelized", "create_pdws", "predict_event", "event_peak_time", "next_event_time", "stft", "spectrogram_my_iq", "plan_time_shards", "stitch_rows", "ChannelizerError", "lib", "LIB_PATH"]
