/* C entry point over the reference's RocJpegStreamParser, compiled unmodified
 * from /root/reference/src/rocjpeg_parser.cpp (see oracle/build.py).
 * ORACLE infrastructure only: used to check that this repository's parser and
 * the oracle's orc_parse make the same accept/reject decisions and extract
 * the same fields. */
#include "rocjpeg_parser.h"
#include <cstring>

struct RefParsed {
    int32_t ok;
    int32_t width, height, ncomp, css;
    int32_t comp_id[3], hs[3], vs[3], tq[3];
    int32_t scan_ncomp, td[3], ta[3];
    int32_t restart_interval;
    uint32_t num_mcus, scan_offset, scan_size;
    uint8_t qt[4][64];
    uint8_t qt_present[4];
    uint8_t dc_bits[2][16], dc_vals[2][12], ac_bits[2][16], ac_vals[2][162];
    uint8_t huff_present[2];
};

extern "C" int ref_parse(const uint8_t *data, uint32_t len, RefParsed *o) {
    std::memset(o, 0, sizeof(*o));
    RocJpegStreamParser parser;
    o->ok = parser.ParseJpegStream(data, len) ? 1 : 0;
    if (!o->ok) return 0;
    const JpegStreamParameters *p = parser.GetJpegStreamParameters();
    o->width = p->picture_parameter_buffer.picture_width;
    o->height = p->picture_parameter_buffer.picture_height;
    o->ncomp = p->picture_parameter_buffer.num_components;
    o->css = p->chroma_subsampling;
    for (int i = 0; i < 3; i++) {
        o->comp_id[i] = p->picture_parameter_buffer.components[i].component_id;
        o->hs[i] = p->picture_parameter_buffer.components[i].h_sampling_factor;
        o->vs[i] = p->picture_parameter_buffer.components[i].v_sampling_factor;
        o->tq[i] = p->picture_parameter_buffer.components[i].quantiser_table_selector;
        o->td[i] = p->slice_parameter_buffer.components[i].dc_table_selector;
        o->ta[i] = p->slice_parameter_buffer.components[i].ac_table_selector;
    }
    o->scan_ncomp = p->slice_parameter_buffer.num_components;
    o->restart_interval = p->slice_parameter_buffer.restart_interval;
    o->num_mcus = p->slice_parameter_buffer.num_mcus;
    o->scan_offset = (uint32_t)(p->slice_data_buffer - data);
    o->scan_size = p->slice_parameter_buffer.slice_data_size;
    for (int t = 0; t < 4; t++) {
        std::memcpy(o->qt[t], p->quantization_matrix_buffer.quantiser_table[t], 64);
        o->qt_present[t] = p->quantization_matrix_buffer.load_quantiser_table[t];
    }
    for (int t = 0; t < 2; t++) {
        std::memcpy(o->dc_bits[t], p->huffman_table_buffer.huffman_table[t].num_dc_codes, 16);
        std::memcpy(o->dc_vals[t], p->huffman_table_buffer.huffman_table[t].dc_values, 12);
        std::memcpy(o->ac_bits[t], p->huffman_table_buffer.huffman_table[t].num_ac_codes, 16);
        std::memcpy(o->ac_vals[t], p->huffman_table_buffer.huffman_table[t].ac_values, 162);
        o->huff_present[t] = p->huffman_table_buffer.load_huffman_table[t];
    }
    return 1;
}
