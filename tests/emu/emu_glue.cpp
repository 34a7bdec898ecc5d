// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Marks the emulated build so that a test can tell which library it loaded.
extern "C" int fi_emu_marker(void) { return 1; }
