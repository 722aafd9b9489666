// pooled_slide_nb.cu -- the sliding-window pooled-histogram kernel of pooled_slide.cu, compiled a second time with
// cooperative 4-byte row stores instead of bulk copies (output rows that are not 16-byte aligned chunks: the 93-channel
// tensor with w % 4 != 0, or an unaligned output pointer).  See the note at the top of pooled_slide.cu.
#define SHDR_SLIDE_NOBULK 1
#include "pooled_slide.cu"
