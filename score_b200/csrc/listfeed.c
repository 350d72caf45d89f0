/* Host-side input marshalling of the drop-in boundary: nested Python lists -> one contiguous int32 array.
 *
 * The reference's loader yields batch_data as nested Python lists (graph_loader.py:383; dummy slices are FLOAT zeros,
 * graph_loader.py:90-91) and TensorFlow converts them inside sess.run (score.py:102-115).  SCORE.train / eval take the
 * same lists; NumPy's generic converter needs ~100 ns per element for them (50 ms for a Taobao batch of 497 K ids - 130 x
 * the device step).  This walks the lists with the list / int / float fast paths of the C API instead.
 *
 * CPython extension, no CUDA, no NumPy headers (the destination is any writable buffer).  Not part of the C ABI of
 * include/score_b200.h: it serves the Python mirror only.
 *
 *   fill_i32(obj, shape, out) -> None
 *     obj    nested lists / tuples of depth len(shape); leaves: int, float (truncated toward zero like ndarray.astype),
 *            or anything with __index__ / __float__ (NumPy scalars)
 *     shape  tuple of ints; every level must have exactly that length (ValueError otherwise, like np.asarray on ragged input)
 *     out    writable C-contiguous buffer of prod(shape) int32
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static int leaf_to_i32(PyObject* o, int32_t* dst) {
    if (PyLong_CheckExact(o)) {
        int overflow = 0;
        long v = PyLong_AsLongAndOverflow(o, &overflow);
        if (overflow || v < INT32_MIN || v > INT32_MAX) {
            PyErr_SetString(PyExc_OverflowError, "id does not fit int32");
            return -1;
        }
        *dst = (int32_t)v;
        return 0;
    }
    if (PyFloat_CheckExact(o)) {
        double d = PyFloat_AS_DOUBLE(o);
        if (!(d > -2147483649.0 && d < 2147483648.0)) {
            PyErr_SetString(PyExc_OverflowError, "id does not fit int32");
            return -1;
        }
        *dst = (int32_t)d;   /* truncation toward zero: ndarray.astype(int32) */
        return 0;
    }
    if (PyList_Check(o) || PyTuple_Check(o)) {
        PyErr_SetString(PyExc_ValueError, "batch_data is nested deeper than its expected shape");
        return -1;
    }
    if (PyIndex_Check(o)) {   /* bool, NumPy integer scalars */
        PyObject* n = PyNumber_Index(o);
        if (!n) return -1;
        int rc = leaf_to_i32(n, dst);
        Py_DECREF(n);
        return rc;
    }
    {
        double d = PyFloat_AsDouble(o);   /* NumPy float scalars, anything with __float__ */
        if (d == -1.0 && PyErr_Occurred()) return -1;
        if (!(d > -2147483649.0 && d < 2147483648.0)) {
            PyErr_SetString(PyExc_OverflowError, "id does not fit int32");
            return -1;
        }
        *dst = (int32_t)d;
        return 0;
    }
}

static int fill_level(PyObject* o, const Py_ssize_t* shape, int depth, int ndim, int32_t** cursor) {
    const Py_ssize_t want = shape[depth];
    PyObject** items;
    Py_ssize_t n;
    if (PyList_CheckExact(o)) {
        n = PyList_GET_SIZE(o);
        items = ((PyListObject*)o)->ob_item;
    } else if (PyTuple_CheckExact(o)) {
        n = PyTuple_GET_SIZE(o);
        items = ((PyTupleObject*)o)->ob_item;
    } else {
        PyErr_Format(PyExc_ValueError, "batch_data: expected a list at depth %d, got %.80s", depth, Py_TYPE(o)->tp_name);
        return -1;
    }
    if (n != want) {
        PyErr_Format(PyExc_ValueError, "batch_data: a list at depth %d has %zd entries, expected %zd", depth, n, want);
        return -1;
    }
    if (depth == ndim - 1) {
        int32_t* dst = *cursor;
        for (Py_ssize_t i = 0; i < n; ++i)
            if (leaf_to_i32(items[i], dst + i) < 0) return -1;
        *cursor = dst + n;
        return 0;
    }
    /* (software prefetch of the next sublists was measured: no gain, the walk is already ~17 ns per id) */
    for (Py_ssize_t i = 0; i < n; ++i)
        if (fill_level(items[i], shape, depth + 1, ndim, cursor) < 0) return -1;
    return 0;
}

static PyObject* fill_i32(PyObject* self, PyObject* args) {
    PyObject *obj, *shape_obj, *out_obj;
    (void)self;
    if (!PyArg_ParseTuple(args, "OO!O", &obj, &PyTuple_Type, &shape_obj, &out_obj)) return NULL;
    const Py_ssize_t ndim = PyTuple_GET_SIZE(shape_obj);
    if (ndim < 1 || ndim > 8) {
        PyErr_SetString(PyExc_ValueError, "shape must have 1..8 dimensions");
        return NULL;
    }
    Py_ssize_t shape[8], total = 1;
    for (Py_ssize_t i = 0; i < ndim; ++i) {
        shape[i] = PyLong_AsSsize_t(PyTuple_GET_ITEM(shape_obj, i));
        if (shape[i] < 0) {
            if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "negative dimension");
            return NULL;
        }
        total *= shape[i];
    }
    Py_buffer view;
    if (PyObject_GetBuffer(out_obj, &view, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (view.len != total * (Py_ssize_t)sizeof(int32_t)) {
        PyBuffer_Release(&view);
        PyErr_SetString(PyExc_ValueError, "output buffer size does not match the shape");
        return NULL;
    }
    int32_t* cursor = (int32_t*)view.buf;
    int rc = 0;
    if (total > 0 || ndim > 0) rc = fill_level(obj, shape, 0, (int)ndim, &cursor);
    PyBuffer_Release(&view);
    if (rc < 0) return NULL;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"fill_i32", fill_i32, METH_VARARGS, "fill_i32(nested_lists, shape, out): flatten into a writable int32 buffer"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_listfeed", "nested-list feed of the SCoRe drop-in boundary", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__listfeed(void) { return PyModule_Create(&module); }
