// idc_swig_common.i -- shared by custom_invlists.swig and altid.swig.
//
// Every wrapped call runs with the GIL released and turns C++ exceptions into Python ones, the way the reference
// modules do (custom_invlists.swig:38-57, altid.swig:34-53): faiss::FaissException and the adapter's
// std::runtime_error (a failed C-ABI call, message from idc_last_error()) -> RuntimeError, std::bad_alloc ->
// MemoryError.
%include <stdint.i>

#define FAISS_API

%exception {
    PyThreadState* idc_saved_ts = PyEval_SaveThread();
    bool idc_failed = false;
    try {
        $action
    } catch (const faiss::FaissException& ex) {
        PyEval_RestoreThread(idc_saved_ts);
        idc_failed = true;
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_RuntimeError, ex.what());
    } catch (const std::bad_alloc&) {
        PyEval_RestoreThread(idc_saved_ts);
        idc_failed = true;
        PyErr_SetString(PyExc_MemoryError, "std::bad_alloc");
    } catch (const std::exception& ex) {
        PyEval_RestoreThread(idc_saved_ts);
        idc_failed = true;
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_RuntimeError, ex.what());
    }
    if (idc_failed) SWIG_fail;
    PyEval_RestoreThread(idc_saved_ts);
}
