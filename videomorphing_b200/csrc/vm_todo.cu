// vm_todo.cu -- entry points of include/vmorph.h that are declared but not built yet in this revision.
// They fail loudly (no silent fallback).
#include "vm_host.h"
using namespace vm;
extern "C" {
int vm_qpath_optimize(int, const float *, float *, int, int, int, float, int *, void *) {
    set_error("vm_qpath_optimize: not built in this revision"); return VM_ERR_STATE; }
int vm_params_parse_xml(const char *, vm_params *, vm_tracks *) { set_error("vm_params_parse_xml: not built in this revision"); return VM_ERR_STATE; }
void vm_tracks_free(vm_tracks *) {}
}
