/* Embeds the sm_100a cubin built from mip_kernels.cu into libfloor_b200_mip.so
   (the counterpart of libfloor #embed-ing etc/mip_map_minify/mmm.fubar, src/device/device_image.cpp:133-139). */
	.section .rodata
	.balign 64
	.global flmip_cubin_begin
	.type flmip_cubin_begin, @object
flmip_cubin_begin:
#ifndef FLMIP_CUBIN_PATH
#define FLMIP_CUBIN_PATH "mip_kernels.cubin"
#endif
	.incbin FLMIP_CUBIN_PATH
	.global flmip_cubin_end
	.type flmip_cubin_end, @object
flmip_cubin_end:
	.byte 0
	.section .note.GNU-stack,"",@progbits
