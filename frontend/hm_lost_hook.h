// Force-included into HM's TDecTop.cpp (frontend/Makefile): the one call the patched TDecTop::xCreateLostPicture makes.
#pragma once
class TComPic;
void hm_hook_lost_picture(TComPic* fill, TComPic* src, int poc);      // frontend/hm_hooks.cpp
