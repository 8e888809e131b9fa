// oracle/standalone_decl.h -- TEST INFRASTRUCTURE: declaration of the reporting hook standalone_hash.sh injects
#pragma once
extern "C" void standalone_note_blend(const unsigned char* p, int w, int h);
