"""tools/fortran_subset.py -- the interpreter that executes the reference's Fortran to make tests/golden/reference_vectors.npz
and reference_global.npz -- on small fixed-form programs written for this test (none of it is reference code): the language
rules the golden vectors depend on.  Each case states the Fortran rule it checks; expected values are worked out by hand."""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fortran_subset as F  # noqa: E402


def run(src, name, *args, consts=None, module_vars=None, members=None, factories=None):
    it = F.Interpreter(consts or {})
    if module_vars:
        it.module_vars.update(module_vars)
    if members:
        it.module_members.update(members)
    if factories:
        it.derived_factories.update(factories)
    it.load(src)
    return it, it.call(name, *args)


def test_operator_precedence_and_association():
    """a / b * c is (a / b) * c; -a ** 2 is -(a ** 2); ** associates to the right; integer division truncates toward zero;
    mixed integer / real division is real"""
    src = """
      subroutine prec( r )
      implicit none
      double precision :: r(8)
      integer :: i, j
      i = 7
      j = 2
      r(1) = 8.0d0 / 4.0d0 * 2.0d0
      r(2) = -3.0d0 ** 2
      r(3) = 2.0d0 ** 3 ** 2
      r(4) = i / j
      r(5) = (-i) / j
      r(6) = i / 2.0d0
      r(7) = 1.0d0 - 2.0d0 - 3.0d0
      r(8) = 2.0d0 * 3.0d0 + 4.0d0 / 8.0d0 - 1.0d0
      return
      end
"""
    r = np.zeros(8)
    run(src, "prec", r)
    assert list(r) == [4.0, -9.0, 512.0, 3.0, -3.0, 3.5, -4.0, 5.5]


def test_do_loops_zero_trip_step_exit_cycle():
    """a do loop whose range is empty runs zero times; negative steps; exit and cycle; do while"""
    src = """
      subroutine loops( n, r )
      implicit none
      integer :: n, i, k
      double precision :: r(5)
      k = 0
      do i = 3, 2
        k = k + 1
      end do
      r(1) = k
      k = 0
      do i = 10, 1, -3
        k = k + i
      end do
      r(2) = k
      k = 0
      do i = 1, n
        if( mod(i,2) .eq. 0 ) cycle
        if( i .gt. 7 ) exit
        k = k + i
      end do
      r(3) = k
      r(4) = i
      k = 0
      do while ( k .lt. 5 )
        k = k + 2
      end do
      r(5) = k
      return
      end
"""
    r = np.zeros(5)
    run(src, "loops", 20, r)
    assert list(r) == [0.0, 10 + 7 + 4 + 1, 1 + 3 + 5 + 7, 9.0, 6.0]


def test_column_major_storage_and_sequence_association():
    """arrays are column major; an element passed as an actual argument hands the callee the storage from that element on,
    seen through the dummy's own shape (explicit or assumed size)"""
    src = """
      subroutine outer( a, r )
      implicit none
      double precision :: a(3,4), r(4)
      call inner( a(2,2), r )
      call tail( a(1,3), r )
      return
      end
      subroutine inner( b, r )
      implicit none
      double precision :: b(2,2), r(4)
      r(1) = b(1,1)
      r(2) = b(2,2)
      b(1,2) = -1.0d0
      return
      end
      subroutine tail( c, r )
      implicit none
      double precision :: c(*), r(4)
      r(3) = c(1)
      r(4) = c(6)
      return
      end
"""
    a = np.asfortranarray(np.arange(1.0, 13.0).reshape(4, 3).T)          # a(i,j) = i + 3 (j-1)
    r = np.zeros(4)
    run(src, "outer", a, r)
    # storage from a(2,2) on: 5, 6, 7, 8 ... -> b(1,1) = 5, b(2,2) = 8, b(1,2) = a(1,3) = 7 is overwritten with -1
    assert list(r) == [5.0, 8.0, -1.0, 12.0] and a[0, 2] == -1.0


def test_sections_vector_subscripts_and_nonconforming_assignment():
    """array sections are views that can be assigned; vector subscripts gather; a longer right-hand side is cut to the extent of
    the left-hand side (what the reference's compiler does with its non-conforming assignments)"""
    src = """
      subroutine sect( a, r )
      implicit none
      double precision :: a(4,4), r(6), w(3)
      integer :: iv(2)
      iv(1) = 4
      iv(2) = 2
      a(2,1:3) = 9.0d0
      r(1:2) = a(iv,4)
      w = a(1:4,2)
      r(3) = w(3)
      r(4) = sum( a(2,1:4) )
      r(5:6) = (/ 1.5d0, 2.5d0 /)
      return
      end
"""
    a = np.asfortranarray(np.arange(1.0, 17.0).reshape(4, 4).T)
    r = np.zeros(6)
    run(src, "sect", a, r)
    assert list(r) == [16.0, 14.0, 7.0, 9.0 * 3 + 14.0, 1.5, 2.5]


def test_data_statements_equivalence_and_select_case():
    """data fills an array in storage order; equivalence( matrix, vector ) shares storage; select case with a default"""
    src = """
      subroutine dat( k, r )
      implicit none
      integer :: k
      double precision :: r(5), m(2,2), v(4), t(3)
      double precision, parameter :: one = 1.0d0, two = 2.0d0
      equivalence ( m, v )
      data t / one, two, 3.0d0 /
      m(1,1) = 1.0d0
      m(2,1) = 2.0d0
      m(1,2) = 3.0d0
      m(2,2) = 4.0d0
      v(3) = 30.0d0
      r(1) = m(1,2)
      r(2) = t(1) + t(2) * t(3)
      select case ( k )
      case ( 1 )
        r(3) = 10.0d0
      case ( 2 )
        r(3) = 20.0d0
      case default
        r(3) = -1.0d0
      end select
      return
      end
"""
    for k, want in ((1, 10.0), (2, 20.0), (5, -1.0)):
        r = np.zeros(5)
        run(src, "dat", k, r)
        assert r[0] == 30.0 and r[1] == 7.0 and r[2] == want


def test_internal_procedures_and_derived_types():
    """an internal procedure shares its host's variables; components of elements of an array of objects (a%b(i,j)%c(k)), as values
    and as the start of a storage run passed to a callee"""
    src = """
      subroutine host( box, r )
      implicit none
      double precision :: r(4), acc
      integer :: i
      acc = 0.0d0
      do i = 1, 3
        call add_one
      end do
      r(1) = acc
      r(2) = box%cells(2,1)%w(3)
      call take( box%cells(1,1)%w(2), r )
      return
      contains
      subroutine add_one
      implicit none
      acc = acc + dble(i)
      return
      end subroutine add_one
      end
      subroutine take( x, r )
      implicit none
      double precision :: x(*), r(4)
      r(3) = x(1) + x(2)
      return
      end
"""
    cells = np.empty((2, 1), dtype=object)
    cells[0, 0] = NS(w=np.array([1.0, 2.0, 3.0]))
    cells[1, 0] = NS(w=np.array([4.0, 5.0, 6.0]))
    r = np.zeros(4)
    run(src, "host", NS(cells=cells), r)
    assert list(r[:3]) == [6.0, 6.0, 5.0]


def test_functions_intrinsics_and_ieee_division():
    """function subprograms; sign, mod, nint, int toward zero, exponent; division by zero gives an IEEE infinity, not an exception"""
    src = """
      double precision function half( x )
      implicit none
      double precision :: x
      half = x / 2.0d0
      return
      end
      subroutine intr( r )
      implicit none
      double precision :: r(8), z
      double precision, external :: half
      z = 0.0d0
      r(1) = half( 9.0d0 )
      r(2) = sign( 3.0d0, -0.5d0 )
      r(3) = mod( -7, 3 )
      r(4) = nint( 2.5d0 )
      r(5) = int( -2.7d0 )
      r(6) = exponent( 8.0d0 )
      r(7) = 1.0d0 / z
      r(8) = dble( 7 / 2 )
      return
      end
"""
    r = np.zeros(8)
    run(src, "intr", r)
    assert list(r[:6]) == [4.5, -3.0, -1.0, 3.0, -2.0, 4.0] and np.isinf(r[6]) and r[7] == 3.0


def test_module_variables_and_library_calls():
    """`use module` brings in the harness's module variables (arrays by reference); BLAS / LAPACK calls: dgemv in the reference BLAS'
    loop order, dgesv through LAPACK with its pivots and info, ddot / dnrm2 on storage runs"""
    src = """
      subroutine lib( x, r )
      use store
      implicit none
      double precision :: x(2), r(5), y(2), aa(2,2), bb(2)
      integer :: ipiv(2), info
      double precision, external :: ddot, dnrm2
      y = 0.0d0
      call dgemv( 'N', 2, 2, 1.0d0, mat, 2, x, 1, 0.0d0, y, 1 )
      r(1) = y(1)
      r(2) = y(2)
      aa = mat
      bb = y
      call dgesv( 2, 1, aa, 2, ipiv, bb, 2, info )
      r(3) = bb(1) - x(1) + bb(2) - x(2) + dble( info )
      r(4) = ddot( 2, mat(1,2), 1, x(1), 1 )
      r(5) = dnrm2( 4, mat(1,1), 1 )
      mat(1,1) = 100.0d0
      return
      end
"""
    mat = np.asfortranarray(np.array([[1.0, 2.0], [3.0, 4.0]]))
    x = np.array([1.0, -1.0])
    r = np.zeros(5)
    run(src, "lib", x, r, module_vars=dict(mat=mat), members=dict(store={"mat"}))
    assert list(r[:2]) == [-1.0, -1.0]
    assert abs(r[2]) < 1e-14 and r[3] == 2.0 * 1.0 + 4.0 * -1.0 and abs(r[4] - np.sqrt(30.0)) < 1e-15
    assert mat[0, 0] == 100.0                                              # the module array itself was written


def test_arithmetic_is_not_contracted():
    """a * b + c is two roundings (no fused multiply-add), in the stated order: the reference's results depend on it only at the
    1e-16 level, the bit-for-bit fixtures (formG, ddot42n) do"""
    src = """
      subroutine fma( a, b, c, r )
      implicit none
      double precision :: a, b, c, r(1)
      r(1) = a * b + c
      return
      end
"""
    a, b = 1.0 + 2.0 ** -30, 1.0 - 2.0 ** -30
    c = -1.0
    r = np.zeros(1)
    run(src, "fma", a, b, c, r)
    assert r[0] == (a * b) + c == 0.0                   # the product rounds to 1 - 2^-60 -> 1.0; a fused operation would give -2^-60


def test_variables_that_begin_with_a_keyword():
    """use_max = .true. is an assignment, not a use statement (nor are save_it, stop_now, format_id, write_flag statements of
    those keywords) -- found when the reference's mm10_set_history_locs came out with the small history layout for 48 systems"""
    src = """
      subroutine kw( r )
      implicit none
      double precision :: r(5)
      logical :: use_max
      integer :: save_it, stop_now, format_id, write_flag
      use_max = .false.
      if( r(1) .gt. 0.0d0 ) then
        use_max = .true.
      end if
      save_it = 2
      stop_now = 3
      format_id = 4
      write_flag = 5
      r(1) = 0.0d0
      if( use_max ) r(1) = 1.0d0
      r(2) = save_it
      r(3) = stop_now
      r(4) = format_id
      r(5) = write_flag
      return
      end
"""
    r = np.ones(5)
    run(src, "kw", r)
    assert list(r) == [1.0, 2.0, 3.0, 4.0, 5.0]


def test_module_scalars_written_by_a_unit_reach_the_harness():
    """a unit that assigns module scalars hands them back; units compiled afterwards see the new values; allocate of a module
    array the harness provides is a shape check"""
    src = """
      subroutine setit
      use sizes, only : ntotal, table
      implicit none
      integer :: i
      if( .not. allocated( table ) ) allocate( table(3) )
      ntotal = 0
      do i = 1, 3
        table(i) = 10 * i
        ntotal = ntotal + table(i)
      end do
      return
      end
      subroutine readit( r )
      use sizes, only : ntotal
      implicit none
      double precision :: r(1)
      r(1) = ntotal
      return
      end
"""
    table = np.zeros(3, dtype=np.int64)
    it, _ = run(src, "setit", module_vars=dict(ntotal=-1, table=table))
    assert it.module_vars["ntotal"] == 60 and list(table) == [10, 20, 30]
    r = np.zeros(1)
    it.call("readit", r)
    assert r[0] == 60.0
